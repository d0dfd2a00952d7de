"""Edge cases of the sweep against the oracle: degenerate problem shapes (one tet, isolated vertices, everything
constrained, no iterations), extreme valence (a vertex with 80 incident tets, more than one warp pass per lane class),
the singular-Hessian skip (sim/vbd/Kernels.h:336-337), disconnected components, ragged colour classes."""
import numpy as np
import pytest

import oracle
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

pytestmark = pytest.mark.gpu
TOL = 1e-4


def rel_l2(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def pair(X, T, *, dbc=None, cheb=None, v=None, detH_zero=None, strategy=None, **tuning):
    d = pbat.sim.vbd.Data().with_volume_mesh(X, T)
    if dbc is not None:
        d = d.with_dirichlet_vertices(dbc)
    if cheb:
        d = d.with_chebyshev_acceleration(cheb)
    if v is not None:
        d = d.with_velocity(v)
    if detH_zero is not None:
        d = d.with_hessian_determinant_zero(detH_zero)
    if strategy is not None:
        d = d.with_initialization_strategy(strategy)
    d = d.construct()
    vbd = pbat.gpu.vbd.Integrator(d, **tuning)
    ref = oracle.Oracle(X, T, dbc=dbc, colors=d.colors, v=v, accel=oracle.ACCEL_CHEBYSHEV if cheb else oracle.ACCEL_NONE,
                        rho=cheb or 1.0, strategy=int(d.strategy), detH_zero=d.detH_zero)
    return d, vbd, ref


def run(vbd, ref, steps, dt=0.01, iters=10, substeps=1):
    for _ in range(steps):
        vbd.step(dt, iters, substeps)
        ref.step(dt, iters, substeps)


def icosphere_star():
    """80 tets sharing the centre vertex: the once-subdivided icosahedron coned to the origin."""
    t = (1 + 5 ** 0.5) / 2
    P = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], float)
    P /= np.linalg.norm(P, axis=1, keepdims=True)
    F = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7),
         (9, 8, 1)]
    pts, mid, F2 = [p for p in P], {}, []

    def m(a, b):
        k = (min(a, b), max(a, b))
        if k not in mid:
            q = pts[a] + pts[b]
            pts.append(q / np.linalg.norm(q))
            mid[k] = len(pts) - 1
        return mid[k]
    for a, b, c in F:
        ab, bc, ca = m(a, b), m(b, c), m(c, a)
        F2 += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
    X = np.concatenate([np.zeros((1, 3)), 0.3 * np.array(pts)]).T           # vertex 0 = centre
    T = np.array([[0, a + 1, b + 1, c + 1] for a, b, c in F2], dtype=np.int64).T
    vol = meshes.tet_volumes(X, T)
    T[:, vol < 0] = T[:, vol < 0][[0, 2, 1, 3]]
    assert (meshes.tet_volumes(X, T) > 0).all() and T.shape[1] == 80
    return X, np.ascontiguousarray(T)


def test_single_tet():
    X = np.array([[0., 1., 0., 0.], [0., 0., 1., 0.], [0., 0., 0., 1.]])
    T = np.array([[0], [1], [2], [3]], dtype=np.int64)
    d, vbd, ref = pair(X, T, dbc=np.array([0]))
    run(vbd, ref, 20)
    assert rel_l2(vbd.x, ref.x) < TOL and rel_l2(vbd.v, ref.v) < 1e-3


def test_isolated_vertices_keep_the_reference_semantics():
    """Vertices no tet references have zero mass: H = 0, |det H| <= detHZero, the update is skipped
    (sim/vbd/Kernels.h:336-337) and they stay wherever the initial guess put them."""
    X, T = meshes.tet_grid(2, 2, 2, 0.5)
    X = np.concatenate([X, [[3.0, 4.0], [3.0, 4.0], [3.0, 4.5]]], axis=1)      # two extra, unreferenced vertices
    v = np.zeros_like(X)
    v[0, -1] = 1.0
    for strategy in (pbat.sim.vbd.InitializationStrategy.Inertia, pbat.sim.vbd.InitializationStrategy.AdaptivePbat):
        d, vbd, ref = pair(X, T, dbc=np.flatnonzero(X[2] == 0), v=v, strategy=strategy)
        run(vbd, ref, 5)
        assert np.isfinite(vbd.x).all()
        assert np.allclose(vbd.x[:, -2:], ref.x[:, -2:], atol=1e-6)
        assert rel_l2(vbd.x, ref.x) < TOL


def test_everything_constrained_and_no_iterations():
    X, T = meshes.tet_grid(3, 2, 2, 0.2)
    d, vbd, ref = pair(X, T, dbc=np.arange(X.shape[1]))
    run(vbd, ref, 3)
    assert np.array_equal(vbd.x, X.astype(np.float32)) and not vbd.v.any()
    # iterations = 0: the step is pre-step + velocity update only
    v = 0.1 * np.random.default_rng(0).standard_normal(X.shape)
    d, vbd, ref = pair(X, T, dbc=np.flatnonzero(X[0] == 0), v=v)
    run(vbd, ref, 2, iters=0, substeps=2)
    assert rel_l2(vbd.x, ref.x) < 1e-6 and np.allclose(vbd.v, ref.v, atol=1e-4)


@pytest.mark.parametrize("tile_iters", [0, 1, 3])
@pytest.mark.parametrize("cheb", [None, 0.85])
def test_extreme_valence(cheb, tile_iters):
    """A vertex with 80 incident tets (the grids have at most 32) next to vertices with 5-6: every lanes-per-vertex
    class, several record blocks per lane, ragged tiles."""
    X, T = icosphere_star()
    X = X + 0.01 * np.random.default_rng(2).uniform(-1, 1, X.shape)
    dbc = np.flatnonzero(X[2] > 0.25)
    d, vbd, ref = pair(X, T, dbc=dbc, cheb=cheb, tile_iters=tile_iters)
    assert np.bincount(T.reshape(-1))[0] == 80
    run(vbd, ref, 15)
    err = rel_l2(vbd.x, ref.x)
    print(f"valence-80 star, cheb={cheb}, tile_iters={tile_iters}: rel L2 = {err:.3e}")
    assert err < TOL


def test_extreme_valence_kernel_variants_agree():
    X, T = icosphere_star()
    dbc = np.flatnonzero(X[2] > 0.25)
    d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_chebyshev_acceleration(0.8).construct()
    out = []
    for variant in (1, 2, 3, 4):
        vbd = pbat.gpu.vbd.Integrator(d, kernel_variant=variant)
        for _ in range(5):
            vbd.step(0.01, 10, 1)
        out.append(vbd.x.copy())
    assert np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[2])


def test_singular_hessian_is_skipped():
    """|det H| <= detHZero => the vertex keeps its position (sim/vbd/Kernels.h:336-337).  With a huge threshold no
    vertex ever moves off the initial guess; with a threshold between the determinants of light and heavy vertices
    only some do -- same set as the oracle's."""
    X, T = meshes.tet_grid(4, 3, 3, 0.1)
    dbc = np.flatnonzero(X[0] == 0)
    d, vbd, ref = pair(X, T, dbc=dbc, detH_zero=1e300)
    run(vbd, ref, 3)
    assert rel_l2(vbd.x, ref.x) < 1e-6
    # determinants here are ~ (m/dt^2 + k)^3 with k ~ 1e5: pick thresholds inside their spread
    for z in (1e14, 1e15, 3e15):
        d, vbd, ref = pair(X, T, dbc=dbc, detH_zero=z)
        run(vbd, ref, 3)
        moved_g = np.abs(vbd.x.astype(np.float64) - X).max(axis=0) > 1e-7
        moved_r = np.abs(ref.x - X).max(axis=0) > 1e-7
        print(f"detHZero={z:g}: {moved_r.sum()} of {X.shape[1]} vertices move")
        assert np.array_equal(moved_g, moved_r)
        assert rel_l2(vbd.x, ref.x) < TOL


def test_disconnected_components_and_ragged_colours():
    """Bodies of very different sizes in one mesh (colour classes from 1 to hundreds of vertices, empty tiles)."""
    Xa, Ta = meshes.tet_grid(6, 5, 4, 0.1)
    Xb = np.array([[2., 2.3, 2., 2.], [0., 0., 0.3, 0.], [0., 0., 0., 0.3]])
    Tb = np.array([[0], [1], [2], [3]], dtype=np.int64)
    Xc, Tc = meshes.tet_grid(1, 1, 1, 0.2, origin=(3.0, 0.0, 0.0))
    X = np.concatenate([Xa, Xb, Xc], axis=1)
    T = np.concatenate([Ta, Tb + Xa.shape[1], Tc + Xa.shape[1] + 4], axis=1)
    dbc = np.flatnonzero(X[2] == 0)
    for cheb in (None, 0.9):
        d, vbd, ref = pair(X, T, dbc=dbc, cheb=cheb)
        run(vbd, ref, 10)
        assert rel_l2(vbd.x, ref.x) < TOL


def test_invalid_problems_are_rejected():
    X, T = meshes.tet_grid(2, 2, 2, 0.5)
    with pytest.raises(ValueError):                                       # inverted element (fem/Jacobian.h:68-80)
        pbat.sim.vbd.Data().with_volume_mesh(X, T[[1, 0, 2, 3]]).construct()
    d = pbat.sim.vbd.Data().with_volume_mesh(X, T).construct()
    d.E = d.E.copy()
    d.E[0, 0] = X.shape[1] + 7                                            # index out of range
    with pytest.raises((ValueError, RuntimeError)):
        pbat.gpu.vbd.Integrator(d)
    d = pbat.sim.vbd.Data().with_volume_mesh(X, T).construct()
    vbd = pbat.gpu.vbd.Integrator(d)
    for bad in ((0.0, 10, 1), (-0.01, 10, 1), (0.01, -1, 1), (0.01, 10, 0)):
        with pytest.raises(ValueError):
            vbd.step(*bad)
    vbd.step(0.01, 1, 1)                                                  # the handle survives the rejected calls
    assert np.isfinite(vbd.x).all()


@pytest.mark.parametrize("ordering", [0, 1, 2])
def test_device_colouring_equals_the_sequential_one(ordering):
    """the greedy colouring on the device.  With the FirstAvailable selection a vertex' colour depends only
    on the neighbours that come earlier in the visiting order, so colouring in rounds reproduces the sequential result
    (graph/Color.h:45-135) vertex for vertex -- all three orderings; a grid, disconnected stacked bodies, a vertex of valence 80
    next to isolated pieces.  LeastUsed, the reference's default, is inherently sequential and refused."""
    G = pbat.graph
    Xs, Ts = icosphere_star()
    Xa, Ta = meshes.tet_grid(3, 3, 2, 0.2, origin=(2.0, 0.0, 0.0))
    Xg, Tg = meshes.tet_grid(14, 11, 9, 0.1)
    Xb, Tb, _ = meshes.stack_bodies(*meshes.tet_grid(5, 4, 3, 0.1), 3)
    cases = [(Xg, Tg), (Xb, Tb), (np.concatenate([Xs, Xa], axis=1), np.concatenate([Ts, Ta + Xs.shape[1]], axis=1))]
    for X, T in cases:
        nV = X.shape[1]
        host = G.mesh_greedy_color(T, nV, ordering, G.GreedyColorSelectionStrategy.FirstAvailable)
        dev = G.mesh_greedy_color(T, nV, ordering, G.GreedyColorSelectionStrategy.FirstAvailable, device=0)
        assert np.array_equal(host, dev)
        for a in range(4):
            for b in range(a + 1, 4):
                assert (dev[T[a]] != dev[T[b]]).all()
    with pytest.raises(NotImplementedError):
        G.mesh_greedy_color(Tg, Xg.shape[1], 2, G.GreedyColorSelectionStrategy.LeastUsed, device=0)
    # through Data.construct, and stepping with it
    d = (pbat.sim.vbd.Data().with_volume_mesh(Xg, Tg).with_dirichlet_vertices(np.flatnonzero(Xg[2] == 0))
         .with_vertex_coloring_strategy(G.GreedyColorOrderingStrategy(ordering), G.GreedyColorSelectionStrategy.FirstAvailable)
         .construct(coloring_device=0))
    assert np.array_equal(d.colors, G.mesh_greedy_color(Tg, Xg.shape[1], ordering, G.GreedyColorSelectionStrategy.FirstAvailable))
    vbd = pbat.gpu.vbd.Integrator(d)
    vbd.step(0.01, 5, 1)
    assert np.isfinite(vbd.x).all()
