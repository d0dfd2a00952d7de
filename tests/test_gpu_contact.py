"""GPU tests of the contact path: LBVH (against the reference's golden topology), active set
(the reference's own detector test) and the contact-aware solve (against the contact oracle)."""
import os

import numpy as np
import pytest

import oracle
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

pytestmark = pytest.mark.gpu


def morton30(c):
    def expand(v):
        v = (v * 0x00010001) & 0xFF0000FF
        v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3
        v = (v * 0x00000005) & 0x49249249
        return v
    q = np.minimum(np.maximum(c * np.float32(1024), np.float32(0)), np.float32(1023)).astype(np.uint32)
    return (expand(q[0]) * 4 + expand(q[1]) * 2 + expand(q[2])).astype(np.uint32)


def test_bvh_golden_topology():
    """gpu/impl/geometry/Bvh.cu:337-410: the 5 tets of the unit cube (identical boxes => duplicate codes)."""
    P, T = meshes.CUBE_P, meshes.CUBE_T
    L = np.stack([P[:, T[:, e]].min(axis=1) for e in range(5)], axis=1)
    U = np.stack([P[:, T[:, e]].max(axis=1) for e in range(5)], axis=1)
    assert (L == 0).all() and (U == 1).all()  # gpu/impl/geometry/Aabb.cu:18-74
    b = pbat.gpu.geometry.build_bvh(L, U, P.min(axis=1), P.max(axis=1))
    assert b["child"].T.tolist() == [[3, 8], [4, 5], [6, 7], [1, 2]]
    assert b["parent"].tolist() == [-1, 3, 3, 0, 1, 1, 2, 2, 0]
    assert b["rightmost"].T.tolist() == [[7, 8], [4, 5], [6, 7], [5, 7]]
    assert b["inds"].tolist() == [0, 1, 2, 3, 4]


@pytest.mark.parametrize("n", [1, 2, 3, 37, 5000, 70001])
def test_bvh_sort_and_boxes(n):
    rng = np.random.default_rng(n)
    c = rng.uniform(0, 1, (3, n)).astype(np.float32)
    c[:, : n // 7] = c[:, [0]]  # many duplicate codes
    h = rng.uniform(0, 0.01, (3, n)).astype(np.float32)
    L, U = c - h, c + h
    b = pbat.gpu.geometry.build_bvh(L, U, [0, 0, 0], [1, 1, 1])
    codes = morton30(np.float32(0.5) * (L + U))
    order = np.argsort(codes, kind="stable")
    assert np.array_equal(b["inds"], order)             # stable: ties keep the box index order
    assert np.array_equal(b["codes"], codes[order])
    if n == 1:
        return
    ni = n - 1
    child, parent = b["child"], b["parent"]
    assert parent[0] == -1 and sorted(child.reshape(-1).tolist()) == list(range(1, 2 * n - 1))
    assert (parent[child[0]] == np.arange(ni)).all() and (parent[child[1]] == np.arange(ni)).all()
    assert np.array_equal(b["lo"][:, ni:], L[:, order]) and np.array_equal(b["hi"][:, ni:], U[:, order])
    assert np.array_equal(b["lo"][:, :ni], np.minimum(b["lo"][:, child[0]], b["lo"][:, child[1]]))
    assert np.array_equal(b["hi"][:, :ni], np.maximum(b["hi"][:, child[0]], b["hi"][:, child[1]]))


def two_tets():
    XT = np.array([[0., 1., 0., 0.1, 0., 1., 0., 0.1],
                   [0., 0., 1., 0.1, 0., 0., 1., 0.1],
                   [0., 0., 0., 1., 1.01, 1.01, 1.01, 2.01]])
    T = np.array([[0, 4], [1, 5], [2, 6], [3, 7]], dtype=np.int64)
    F = np.array([[0, 1, 2, 0, 4, 5, 6, 4], [1, 2, 0, 2, 5, 6, 4, 6], [3, 3, 3, 1, 7, 7, 7, 5]], dtype=np.int64)
    B = np.array([0, 0, 0, 0, 1, 1, 1, 1])
    v = np.zeros((3, 8))
    v[2, :4], v[2, 4:] = 1.0, -1.0  # dt * v = +-0.01: the motion of the reference's test
    return XT, T, F, B, v


def test_active_set_two_tets():
    """gpu/impl/contact/VertexTriangleMixedCcdDcd.cu:232-329: exactly 4 active vertices (the bottom tet's
    apex and the top tet's bottom face), vertex 3 among them; active vertices have >= 1 neighbour."""
    XT, T, F, B, v = two_tets()
    d = (pbat.sim.vbd.Data().with_volume_mesh(XT, T).with_surface_mesh(np.arange(8), F).with_bodies(B)
         .with_velocity(v).with_acceleration(np.zeros((3, 8)))
         .with_initialization_strategy(pbat.sim.vbd.InitializationStrategy.KineticEnergyMinimum).construct())
    vbd = pbat.gpu.vbd.Integrator(d)
    vbd.scene_bounding_box = (np.array([0., 0., 0.]), np.array([1., 1., 2.01]))
    vbd.step(0.01, 0, 1)  # no solve: detection on xt -> xt + dt v, neighbours on the initial guess
    active, nn, n_active = vbd.contact_state()
    assert n_active == 4
    has_nn = (nn >= 0).any(axis=1)
    assert has_nn.sum() == 4 and has_nn[3] and has_nn[[4, 5, 6]].all()
    ref = oracle.Oracle(XT, T, v=v, aext=np.zeros((3, 8)), colors=d.colors, strategy=2, B=B, V=np.arange(8), F=F)
    ref.step(0.01, 0, 1)
    rnn = ref.get("nn")
    assert np.array_equal(np.sort(nn, axis=1), np.sort(rnn, axis=1))
    assert np.array_equal(active, ref.get("active").astype(bool))


def stacked_scene(n_top=2):
    Xb, Tb = meshes.tet_grid(3, 3, 2, 0.25)
    Xt, Tt = meshes.tet_grid(n_top, n_top, 2, 0.25, origin=(0.1, 0.13, 0.52))
    X = np.concatenate([Xb, Xt], axis=1)
    T = np.concatenate([Tb, Tt + Xb.shape[1]], axis=1)
    B = np.concatenate([np.zeros(Xb.shape[1], np.int64), np.ones(Xt.shape[1], np.int64)])
    F = meshes.boundary_facets(T)
    V = np.unique(F)
    dbc = np.flatnonzero(X[2] == 0)
    v = np.zeros_like(X)
    v[2, Xb.shape[1]:] = -0.5
    return X, T, B, F, V, dbc, v


@pytest.mark.parametrize("cheb", [None, 0.7])
def test_contact_solve_matches_oracle(cheb):
    """A small body dropped onto a fixed-base body: trajectories with active contact vs the double-precision
    contact oracle (brute-force detection, reference contact energy semantics)."""
    X, T, B, F, V, dbc, v = stacked_scene()
    d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_surface_mesh(V, F).with_bodies(B).with_velocity(v)
         .with_dirichlet_vertices(dbc).with_contact_parameters(1e5, 0.3, 1e-3))
    if cheb:
        d = d.with_chebyshev_acceleration(cheb)
    d = d.construct()
    vbd = pbat.gpu.vbd.Integrator(d)
    ref = oracle.Oracle(X, T, v=v, dbc=dbc, colors=d.colors, B=B, V=V, F=F, muC=1e5, muF=0.3, epsv=1e-3,
                        accel=oracle.ACCEL_CHEBYSHEV if cheb else oracle.ACCEL_NONE, rho=cheb or 1.0)
    touched = False
    for s in range(40):
        vbd.step(0.01, 10, 1)
        ref.step(0.01, 10, 1)
        _, nn, _ = vbd.contact_state()
        touched |= bool((nn >= 0).any())
    assert touched, "the scene never produced a contact"
    xr = ref.x
    err = np.linalg.norm(vbd.x - xr) / np.linalg.norm(xr)
    print(f"contact scene cheb={cheb}: rel L2 = {err:.3e}; top body lowest z = {xr[2, B == 1].min():.4f}")
    assert err < 1e-4
    # the top body must be resting on the bottom one, not passing through it
    assert xr[2, B == 1].min() > 0.45


def test_config3_replica_against_oracle():
    """BASELINE configs[2] scaled down (SURVEY.md 8d "C3": bodies stacked in z with a gap of 0.1 x extent, body ids 0.., the
    bottom body's lowest 1 % fixed, muC = 1e6, muF = 0.3, epsv = 1e-3, active-set frequency 1, dt = 0.01, 20 iterations,
    50 steps): three bodies of 6^3 cubes instead of sixteen of 29^3, against the contact oracle.

    Every body is shifted sideways by a generic fraction of a cell: with identical grids exactly above each other every
    surface vertex projects onto a VERTEX or an EDGE of the triangle below, and whether such a pair responds (barycentric
    coordinates in [0, 1], sim/vbd/Kernels.h:253-258) is decided by the last bit -- fp32 and double then disagree at first
    touch (tools/contact_diag.py: 85 = 85 contacts, different lists, 2.5e-5 -> 2.6e-2).  In generic position the two agree on
    every contact list for the first 30 steps (free fall, first impacts, 146 vertices in contact) with positions equal to
    6e-6; the stiff penalty (k = 1e6 x area) then amplifies the first borderline decision, so beyond that only the outcome
    is compared."""
    n = 6
    Xb, Tb = meshes.tet_grid(n, n, n, 1.0 / n)
    X, T, B = meshes.stack_bodies(Xb, Tb, 3, axis=2, gap_frac=0.1)
    for b in range(3):
        X[0, B == b] += 0.37 * b / n
        X[1, B == b] += 0.21 * b / n
    F = meshes.boundary_facets(T)
    V = np.unique(F)
    dbc = np.flatnonzero(X[2] <= X[2].min() + 0.01)
    d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_surface_mesh(V, F).with_bodies(B)
         .with_dirichlet_vertices(dbc).with_contact_parameters(1e6, 0.3, 1e-3).construct())
    vbd = pbat.gpu.vbd.Integrator(d)
    ref = oracle.Oracle(X, T, dbc=dbc, colors=d.colors, B=B, V=V, F=F, muC=1e6, muF=0.3, epsv=1e-3)
    contacts = 0
    for s in range(50):
        vbd.step(0.01, 20, 1)
        ref.step(0.01, 20, 1)
        act, nn, _ = vbd.contact_state()
        if s < 30:
            rnn = ref.get("nn").reshape(-1, 8)
            assert np.array_equal(np.sort(nn, axis=1), np.sort(rnn, axis=1)), f"contact lists differ at step {s}"
            assert np.array_equal(act, ref.get("active").astype(bool))
            err = np.linalg.norm(vbd.x - ref.x) / np.linalg.norm(ref.x)
            assert err < 1e-4, (s, err)
            contacts = max(contacts, int((nn >= 0).any(axis=1).sum()))
    assert contacts > 100, "the bodies never touched"
    xr, xg = ref.x, vbd.x
    err = np.linalg.norm(xg - xr) / np.linalg.norm(xr)
    print(f"config 3 replica: vertices in contact within 30 steps = {contacts}; rel L2 after 50 steps = {err:.3e}")
    assert np.isfinite(xg).all() and err < 3e-2
    for x in (xr, xg):   # same outcome: the stack stays ordered, nothing has passed through (a body is 6 cells high)
        lo = np.array([x[2, B == b].min() for b in range(3)])
        hi = np.array([x[2, B == b].max() for b in range(3)])
        assert (np.diff(lo) > 0).all() and (lo[1:] - hi[:-1] > -3.0 / n).all()


@pytest.mark.parametrize("cheb,substeps,kd", [(None, 1, 0.0), (0.8, 1, 0.0), (None, 2, 1e-4), (0.7, 3, 0.0)])
def test_contact_barrier_free_sweep_is_bit_identical_to_the_barrier_sweep(cheb, substeps, kd, monkeypatch):
    """With contact the whole-GPU kernel sweeps barrier-free too: triangle corners (vertices of OTHER bodies, whose tiles
    may be a sweep ahead or behind) are read from the history of every vertex' last four writes, by write number.  Same
    values read, same arithmetic: identical bits to the sweep with colour barriers (VBDX_DATAFLOW=0), contact lists
    included, over impacts, sliding and several substeps."""
    n = 8
    Xb, Tb = meshes.tet_grid(n, n, n, 1.0 / n)
    X, T, B = meshes.stack_bodies(Xb, Tb, 4, axis=2, gap_frac=0.05)
    for b in range(4):
        X[0, B == b] += 0.37 * b / n
        X[1, B == b] += 0.21 * b / n
    F = meshes.boundary_facets(T)
    V = np.unique(F)
    dbc = np.flatnonzero(X[2] <= X[2].min() + 0.01)
    v = np.zeros_like(X)
    v[2] = -0.5 * B
    d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_surface_mesh(V, F).with_bodies(B).with_velocity(v)
         .with_dirichlet_vertices(dbc).with_contact_parameters(1e5, 0.3, 1e-3).with_rayleigh_damping(kd))
    if cheb:
        d = d.with_chebyshev_acceleration(cheb)
    d = d.construct()
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("VBDX_DATAFLOW", mode)
        vbd = pbat.gpu.vbd.Integrator(d, kernel_variant=3)
        contacts = 0
        for s in range(40):
            vbd.step(0.01, 8, substeps)
            contacts = max(contacts, int((vbd.contact_state()[1] >= 0).any(axis=1).sum()))
        out[mode] = (vbd.x.copy(), vbd.v.copy(), vbd.contact_state(), contacts, vbd.info["blockThreads"])
    assert out["1"][3] > 50, "the bodies never touched"
    assert out["1"][4] != out["0"][4]                     # the lean kernel (no barrier warp) did run
    assert np.isfinite(out["1"][0]).all()
    assert np.array_equal(out["1"][0], out["0"][0]) and np.array_equal(out["1"][1], out["0"][1])
    for a, b in zip(out["1"][2], out["0"][2]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("accel", ["anderson", "broyden"])
def test_windowed_accelerators_with_contact(accel):
    """The reference's example (python/examples/vbd.py:297-330) combines a surface mesh with Anderson / Broyden
    acceleration: its GPU accelerators call the same contact-aware sweep.  Same scene as above against the oracle;
    tolerance of the windowed accelerators (tests/test_gpu_parity.py: 1.5e-3 after the window has mixed fp32 iterates)."""
    X, T, B, F, V, dbc, v = stacked_scene()
    d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_surface_mesh(V, F).with_bodies(B).with_velocity(v)
         .with_dirichlet_vertices(dbc).with_contact_parameters(1e5, 0.3, 1e-3))
    d = (d.with_anderson_acceleration(4) if accel == "anderson" else d.with_broyden_acceleration(4)).construct()
    vbd = pbat.gpu.vbd.Integrator(d)
    ref = oracle.Oracle(X, T, v=v, dbc=dbc, colors=d.colors, B=B, V=V, F=F, muC=1e5, muF=0.3, epsv=1e-3)
    ref.set_acceleration(oracle.ACCEL_ANDERSON if accel == "anderson" else oracle.ACCEL_BROYDEN, window=4)
    touched = False
    for s in range(40):
        vbd.step(0.01, 10, 1)
        ref.step(0.01, 10, 1)
        _, nn, _ = vbd.contact_state()
        touched |= bool((nn >= 0).any())
    assert touched, "the scene never produced a contact"
    xr = ref.x
    err = np.linalg.norm(vbd.x - xr) / np.linalg.norm(xr)
    print(f"contact scene {accel}: rel L2 = {err:.3e}; top body lowest z = {xr[2, B == 1].min():.4f}")
    assert err < 1.5e-3
    assert xr[2, B == 1].min() > 0.45 and vbd.x[2, B == 1].min() > 0.45


def _contact_ref():
    """The reference's own contact code compiled for the GPU (oracle/contact_ref.cu -> oracle/_ref/libcontact_ref.so)."""
    import ctypes as C

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libcontact_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libcontact_ref.so has not been built (needs the reference tree: make -C oracle contact_ref)")
    lib = C.CDLL(path)
    lib.contact_ref_pairs.restype = C.c_int
    lib.contact_ref_pairs.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    lib.contact_ref_penalties.restype = C.c_int
    lib.contact_ref_penalties.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
    return lib


def test_contact_term_against_the_reference_function():
    """csrc/contact.cuh vs pbat::sim::vbd::kernels::AccumulateVertexTriangleContact (sim/vbd/Kernels.h:223-302) itself, both
    run on the GPU in fp32 on the same (vertex, triangle) pairs: random penetrating / separated / sliding / sticking pairs,
    projections outside the triangle, on an edge and on a corner, zero distance, degenerate triangles."""
    from physicsbasedanimationtoolkit_b200 import _lib

    ref = _contact_ref()
    rng = np.random.default_rng(7)
    rows = []

    def pair(xtv, xv, xtf, xf, dt=0.01, k=1e4, muF=0.3, epsv=1e-3):
        rows.append(np.concatenate([xtv, xv, np.asarray(xtf).reshape(-1), np.asarray(xf).reshape(-1), [dt, k, muF, epsv]]))

    for _ in range(4000):
        tri = rng.uniform(-1, 1, (3, 3))                                  # rows = triangle vertices
        n = np.cross(tri[1] - tri[0], tri[2] - tri[0])
        if np.linalg.norm(n) < 0.2:
            continue
        n /= np.linalg.norm(n)
        b = rng.dirichlet([1, 1, 1]) if rng.random() < 0.8 else rng.uniform(-0.5, 1.5, 3)   # 20 %: projection may fall outside
        b = b / b.sum()
        # depth and relative motion well above fp32 resolution of O(1) coordinates: the term differences its inputs
        # ((x_v - x_b) . n, (x_v - x_v^t) - (x_b - x_b^t)), smaller values would only measure cancellation noise -- the
        # stick regime |u| < epsv dt is covered by the exactly representable cases below and by epsv up to 10 here
        depth = rng.choice([-1, 1]) * 10.0 ** rng.uniform(-2.5, -0.5)     # penetrating (< 0) or separated (> 0)
        xv = b @ tri + depth * n
        slide = 10.0 ** rng.uniform(-2.5, -1) * rng.normal(size=3)
        tri_t = tri - 10.0 ** rng.uniform(-3, -1.5) * rng.normal(size=(3, 3))
        pair(xv - slide, xv, tri_t, tri, k=10.0 ** rng.uniform(2, 6), muF=rng.uniform(0, 1), epsv=10.0 ** rng.uniform(-3, 1))
    A, B, Cc = np.array([0., 0, 0]), np.array([1., 0, 0]), np.array([0., 1, 0])
    T0 = [A, B, Cc]
    up = np.array([0., 0, 1])
    tiny = 2.0 ** -20                                                      # below epsv dt = 1e-5: sticking
    for xv in ([0.25, 0.25, -0.125], [0.25, 0.25, 0.0], [0.5, 0.0, -0.25], [0.0, 0.0, -0.25], [0.5, 0.5, -0.5], [1.0, 0.0, -0.5],
               [0.25, 0.25, 0.5], [1.5, 0.25, -0.25], [-0.25, 0.25, -0.25], [0.75, 0.75, -0.25]):
        xv = np.array(xv)
        pair(xv + 2.0 ** -10 * up, xv, T0, T0)                             # moved straight down: no tangential motion
        pair(xv - np.array([2.0 ** -8, 2.0 ** -9, 0.0]), xv, T0, T0)      # slid along the triangle (slipping)
        pair(xv - np.array([tiny, 0.0, 0.0]), xv, T0, T0)                 # ... by less than the threshold (sticking)
        pair(xv, xv, [A, B, Cc + np.array([2.0 ** -9, 0, 0])], T0)         # the triangle moved instead
    pair(np.array([0.1, 0.1, -0.1]), np.array([0.1, 0.1, -0.1]), [A, A, A], [A, A + 1e-12, A])   # degenerate triangle
    pair(np.array([0.1, 0.1, -0.1]), np.array([0.1, 0.1, -0.1]), [A, B, 2 * B], [A, B, 2 * B])   # collinear
    X = np.ascontiguousarray(np.stack(rows), dtype=np.float32)
    n = X.shape[0]
    want = np.zeros((n, 13), np.float32)
    got = np.zeros((n, 13), np.float32)
    assert ref.contact_ref_pairs(n, X.ctypes.data, want.ctypes.data) == 0
    _lib.check(_lib.lib().vbdx_debug_contact_pairs(n, X.ctypes.data, got.ctypes.data))
    assert np.isfinite(want).all() and np.isfinite(got).all()
    # the two agree on WHICH pairs respond (a pair whose projection is within rounding of the triangle's rim may differ:
    # none of the random ones may, the constructed rim cases are exact in binary)
    resp_w, resp_g = np.abs(want[:, 4:]).sum(axis=1) > 0, np.abs(got[:, 4:]).sum(axis=1) > 0
    assert np.array_equal(resp_w, resp_g), np.flatnonzero(resp_w != resp_g)
    assert resp_w.sum() > 2500 and (~resp_w).sum() > 300
    scale_g = np.abs(want[:, 1:4]).max(axis=1, keepdims=True) + 1e-30
    scale_h = np.abs(want[:, 4:]).max(axis=1, keepdims=True) + 1e-30
    err_g = (np.abs(got[:, 1:4] - want[:, 1:4]) / scale_g)[resp_w].max()
    err_h = (np.abs(got[:, 4:] - want[:, 4:]) / scale_h)[resp_w].max()
    # fp32 on both sides, different operation order (expression templates vs scalar code)
    assert err_h < 2e-4 and err_g < 2e-4, (err_g, err_h)
    # gradient direction and magnitude on all responding pairs in aggregate
    assert np.linalg.norm(got[resp_w, 1:4] - want[resp_w, 1:4]) / np.linalg.norm(want[resp_w, 1:4]) < 1e-5


def test_contact_penalty_against_the_reference_struct():
    """The area-scaled penalties of the sweep vs pbat::gpu::impl::vbd::kernels::ContactPenalty<8> (gpu/impl/vbd/Kernels.cuh:80-114)."""
    from physicsbasedanimationtoolkit_b200 import _lib

    ref = _contact_ref()
    rng = np.random.default_rng(3)
    nv, nf = 500, 64
    fc = np.full((nv, 8), -1, np.int32)
    for i in range(nv):
        k = rng.integers(0, 9)
        fc[i, :k] = rng.integers(0, nf, k)
    XVA = rng.uniform(0.01, 1.0, nv).astype(np.float32)
    FA = rng.uniform(0.01, 1.0, nf).astype(np.float32)
    out = {}
    for name in ("ref", "ours"):
        nc, pen = np.zeros(nv, np.int32), np.zeros((nv, 8), np.float32)
        if name == "ref":
            assert ref.contact_ref_penalties(nv, nf, fc.ctypes.data, XVA.ctypes.data, FA.ctypes.data, 1e6, nc.ctypes.data, pen.ctypes.data) == 0
        else:
            _lib.check(_lib.lib().vbdx_debug_contact_penalties(nv, nf, fc.ctypes.data, XVA.ctypes.data, FA.ctypes.data, 1e6, nc.ctypes.data, pen.ctypes.data))
        out[name] = (nc, pen)
    assert np.array_equal(out["ref"][0], out["ours"][0]) and np.array_equal(out["ref"][0], (fc >= 0).sum(axis=1))
    assert np.allclose(out["ref"][1], out["ours"][1], rtol=2e-6, atol=0)
