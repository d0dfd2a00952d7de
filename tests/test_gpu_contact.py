"""GPU tests of the contact path: LBVH (against the reference's golden topology), active set
(the reference's own detector test) and the contact-aware solve (against the contact oracle)."""
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
