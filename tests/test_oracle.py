"""CPU tests of the oracle (test infrastructure): the restated port against the reference's own
known-answer tests and against golden trajectories generated from the reference-header build."""
import os

import numpy as np
import pytest

import oracle
from physicsbasedanimationtoolkit_b200 import meshes

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "vbd_golden.npz"))
from golden.make_golden import CASES, run  # noqa: E402

KINDS = ["port"] + (["reference"] if oracle.have_ref() else [])


@pytest.mark.parametrize("kind", KINDS)
def test_cube_free_fall_known_answer(kind):
    """sim/vbd/Integrator.cpp:245-293: dz < 0, |dxy| < 1e-4, |grad f| < 1e-4, f < f0 (analytic dz = -9.81e-4)."""
    o = oracle.Oracle(meshes.CUBE_P, meshes.CUBE_T, kind=kind)
    assert list(o.get("colors")) == [0, 3, 2, 1, 1, 2, 3, 0]
    dt = 1e-2
    x0 = o.x
    xtilde = x0 + dt * o.v + dt * dt * o.get("aext")
    f0 = o.objective(x0, xtilde, dt)
    o.step(dt, 10, 1)
    dx = o.x - meshes.CUBE_P
    assert (dx[2] < 0).all() and (np.abs(dx[:2]) < 1e-4).all()
    assert np.allclose(dx[2], -9.81e-4, atol=1e-8)
    assert np.linalg.norm(o.objective_gradient(o.x, xtilde, dt)) < 1e-4
    assert o.objective(o.x, xtilde, dt) < f0


@pytest.mark.parametrize("kind", KINDS)
def test_cube_chebyshev_known_answer(kind):
    """sim/vbd/ChebyshevIntegrator.cpp:40-85 (rho = 0.9)"""
    o = oracle.Oracle(meshes.CUBE_P, meshes.CUBE_T, accel=oracle.ACCEL_CHEBYSHEV, rho=0.9, kind=kind)
    o.step(1e-2, 10, 1)
    dx = o.x - meshes.CUBE_P
    assert (dx[2] < 0).all() and (np.abs(dx[:2]) < 1e-4).all()


@pytest.mark.parametrize("kind", KINDS)
def test_cube_anderson_known_answer(kind):
    """sim/vbd/AndersonIntegrator.cpp:62-117 (m = 5): dz < 0, |dxy| < 1e-4, |grad f| < 1e-4, f < f0 after one
    step of 10 iterations.  (The twin doctest of NesterovIntegrator.cpp:50-103 is NOT satisfied by a literal
    restatement of NesterovIntegrator::Solve -- x^{k-1} is never advanced there and the vertices end up above
    their start -- so that accelerator stays unpinned and is not offered by the product)"""
    o = oracle.Oracle(meshes.CUBE_P, meshes.CUBE_T, kind=kind)
    o.set_acceleration(oracle.ACCEL_ANDERSON, window=5)
    dt = 1e-2
    x0 = o.x
    xtilde = x0 + dt * o.v + dt * dt * o.get("aext")
    f0 = o.objective(x0, xtilde, dt)
    o.step(dt, 10, 1)
    dx = o.x - meshes.CUBE_P
    assert (dx[2] < 0).all() and (np.abs(dx[:2]) < 1e-4).all()
    assert np.linalg.norm(o.objective_gradient(o.x, xtilde, dt)) < 1e-4
    assert o.objective(o.x, xtilde, dt) < f0


@pytest.mark.parametrize("kind", KINDS)
def test_cube_broyden_known_answer(kind):
    """sim/vbd/BroydenIntegrator.cpp:83-135 (m = 5, 15 iterations): dz < 0, |dxy| < 1e-4, |grad f| < 1e-4, f < f0."""
    o = oracle.Oracle(meshes.CUBE_P, meshes.CUBE_T, kind=kind)
    o.set_acceleration(oracle.ACCEL_BROYDEN, window=5)
    dt = 1e-2
    x0 = o.x
    xtilde = x0 + dt * o.v + dt * dt * o.get("aext")
    f0 = o.objective(x0, xtilde, dt)
    o.step(dt, 15, 1)
    dx = o.x - meshes.CUBE_P
    assert (dx[2] < 0).all() and (np.abs(dx[:2]) < 1e-4).all()
    assert np.linalg.norm(o.objective_gradient(o.x, xtilde, dt)) < 1e-4
    assert o.objective(o.x, xtilde, dt) < f0


def test_broyden_converges_faster_than_plain_sweeps():
    X, T = meshes.tet_grid(6, 3, 3, 0.1)
    dbc = np.flatnonzero(X[0] == 0)
    plain, acc, conv = (oracle.Oracle(X, T, dbc=dbc) for _ in range(3))
    acc.set_acceleration(oracle.ACCEL_BROYDEN, window=5)
    for _ in range(5):
        plain.step(0.01, 10, 1), acc.step(0.01, 10, 1), conv.step(0.01, 300, 1)
    assert np.linalg.norm(acc.x - conv.x) < 0.25 * np.linalg.norm(plain.x - conv.x)


def test_anderson_least_squares_is_numpy_lstsq():
    """The Anderson mixing weights are Eigen's CompleteOrthogonalDecomposition solve (minimum-norm least
    squares); the oracle's restatement must agree with LAPACK's on a beam where the window fills up."""
    X, T = meshes.tet_grid(4, 2, 2, 0.25)
    dbc = np.flatnonzero(X[0] == 0)
    a = oracle.Oracle(X, T, dbc=dbc)
    a.set_acceleration(oracle.ACCEL_ANDERSON, window=3)
    b = oracle.Oracle(X, T, dbc=dbc, colors=a.get("colors"))     # plain sweeps, Anderson driven from numpy
    dt, iters, m = 1e-2, 8, 3
    a.step(dt, iters, 1)
    # b: pre-step through a zero-iteration step is not available, so replay the step by hand
    xt, v, aext = b.x, b.v, b.get("aext")
    b.step(dt, 0, 1)                                             # sets xt, xtilde and the initial guess, no sweeps
    b.v = v                                                      # (a zero-iteration step also rewrote v)
    n3 = X.size
    flat = lambda M: M.T.reshape(-1)
    DF, DG = np.zeros((n3, m)), np.zeros((n3, m))
    xkm1 = flat(b.x)
    b.sweeps(dt, 1)
    Gkm1 = flat(b.x)
    Fkm1 = Gkm1 - xkm1
    for k in range(1, iters):
        xkm1 = flat(b.x)
        b.sweeps(dt, 1)
        Gk = flat(b.x)
        Fk = Gk - xkm1
        DG[:, (k - 1) % m] = Gk - Gkm1
        DF[:, (k - 1) % m] = Fk - Fkm1
        Gkm1, Fkm1 = Gk, Fk
        mk = min(m, k)
        alpha = np.linalg.lstsq(DF[:, :mk], Fk, rcond=1e-10)[0]
        b.x = (Gk - DG[:, :mk] @ alpha).reshape(-1, 3).T
    assert np.allclose(a.x, b.x, rtol=0, atol=1e-9 * np.abs(a.x).max())


@pytest.mark.parametrize("kind", KINDS)
def test_snh_energy_at_identity(kind):
    """physics/StableNeoHookeanEnergy.cpp:10-43: psi(I) = 0.5 mu (I2 - 3) + 0.5 lambda (I3 - gamma)^2 >= 0"""
    mu, lam = 3.4e5, 3.1e6
    psi = oracle.snh_eval(np.eye(3), mu, lam, kind=kind)
    gamma = 1 + mu / lam
    assert psi >= 0 and abs(psi - 0.5 * lam * (1 - gamma) ** 2) < 1e-9 * max(1.0, psi)


@pytest.mark.parametrize("kind", KINDS)
def test_stvk_energy_known_answer(kind):
    """physics/SaintVenantKirchhoffEnergy.cpp:9-39: psi(I) = 0 and psi(F) = mu |E|^2 + lambda/2 tr(E)^2."""
    mu, lam = 3.4e5, 3.1e6
    assert abs(oracle.stvk_eval(np.eye(3), mu, lam, kind=kind)) <= 1e-15
    F = np.eye(3) + 0.2 * np.random.default_rng(3).standard_normal((3, 3))
    E = 0.5 * (F.T @ F - np.eye(3))
    expected = mu * (E * E).sum() + 0.5 * lam * np.trace(E) ** 2
    assert abs(oracle.stvk_eval(F, mu, lam, kind=kind) - expected) <= 1e-12 * expected


@pytest.mark.parametrize("material", [oracle.MATERIAL_STABLE_NEO_HOOKEAN, oracle.MATERIAL_STVK])
def test_objective_gradient_is_the_derivative_of_the_objective(material):
    """sim/vbd/Integrator.cpp:138-200 restated: central differences of f against the analytic gradient, and the
    per-vertex Newton blocks of the sweep are consistent with it (one sweep from a perturbed state lowers f)."""
    X, T = meshes.tet_grid(3, 2, 2, 0.3)
    o = oracle.Oracle(X, T, material=material)
    rng = np.random.default_rng(1)
    xk = X + 0.02 * rng.standard_normal(X.shape)
    xt = X + 0.01 * rng.standard_normal(X.shape)
    g = o.objective_gradient(xk, xt, 0.05).reshape(-1, 3).T
    for _ in range(12):
        i, dd = rng.integers(X.shape[1]), rng.integers(3)
        h = 1e-6
        xp, xm = xk.copy(), xk.copy()
        xp[dd, i] += h
        xm[dd, i] -= h
        fd = (o.objective(xp, xt, 0.05) - o.objective(xm, xt, 0.05)) / (2 * h)
        assert abs(fd - g[dd, i]) <= 1e-6 * max(1.0, abs(g[dd, i]))
    o.step(0.01, 0, 1)                      # sets xtilde, no sweeps
    xtilde = o.get("xtilde")
    f0 = o.objective(o.x, xtilde, 0.01)
    o.sweeps(0.01, 5)
    assert o.objective(o.x, xtilde, 0.01) < f0


def test_coloring_is_proper_for_all_strategies():
    """graph/Color.cpp:9-49"""
    X, T = meshes.tet_grid(4, 3, 3)
    for ordering in range(3):
        for selection in range(2):
            o = oracle.Oracle(X, T, ordering=ordering, selection=selection)
            c = o.get("colors")
            for a in range(4):
                for b in range(a + 1, 4):
                    assert (c[T[a]] != c[T[b]]).all()


def test_setup_quantities():
    X, T = meshes.tet_grid(3, 2, 2, 0.5)
    o = oracle.Oracle(X, T)
    assert np.allclose(o.get("wg"), meshes.tet_volumes(X, T))
    assert np.isclose(o.get("m").sum(), 1e3 * 3 * 2 * 2 * 0.125)
    GP = o.get("GP")
    assert np.allclose(GP.sum(axis=0), 0, atol=1e-12)  # gradients of a partition of unity
    p, e, il = o.get("GVGp"), o.get("GVGe"), o.get("GVGilocal")
    for i in range(X.shape[1]):
        row = e[p[i]:p[i + 1]]
        assert (np.diff(row) > 0).all() and (T[il[p[i]:p[i + 1]], row] == i).all()


def test_invalid_inputs():
    X, T = meshes.tet_grid(2, 2, 2)
    with pytest.raises(ValueError):
        oracle.Oracle(X, T[[1, 0, 2, 3]])  # inverted tets
    with pytest.raises(ValueError):
        oracle.Oracle(X, T, accel=oracle.ACCEL_CHEBYSHEV, rho=1.5)


@pytest.mark.parametrize("name", [n for n in CASES if not n.startswith("config1")])
def test_port_matches_reference_golden(name):
    """The restated arithmetic (closed-form SNH block) against trajectories produced by the reference's
    own headers (9x9 Hessian route): agreement to round-off."""
    x, v, c = run(name, "port")
    assert np.array_equal(c, GOLD[name + "/colors"])
    assert np.allclose(x, GOLD[name + "/x"], rtol=0, atol=1e-10)
    assert np.allclose(v, GOLD[name + "/v"], rtol=0, atol=1e-8)


def test_port_matches_reference_golden_config1():
    """BASELINE.json configs[0] (100 steps x 20 iterations, ~10k tets)."""
    x, v, c = run("config1_base", "port")
    g = GOLD["config1_base/x"]
    assert np.linalg.norm(x - g) / np.linalg.norm(g) < 1e-9


@pytest.mark.skipif(not oracle.have_ref(), reason="needs oracle/_ref (reference tree)")
def test_live_reference_matches_golden():
    x, _, _ = run("beam_small_cheb", "reference")
    assert np.array_equal(x, GOLD["beam_small_cheb/x"])


def _stacked():
    Xb, Tb = meshes.tet_grid(2, 2, 1, 0.5)
    Xt, Tt = meshes.tet_grid(1, 1, 1, 0.5, origin=(0.2, 0.3, 0.53))
    X = np.concatenate([Xb, Xt], axis=1)
    T = np.concatenate([Tb, Tt + Xb.shape[1]], axis=1)
    B = np.concatenate([np.zeros(Xb.shape[1], np.int64), np.ones(Xt.shape[1], np.int64)])
    F = meshes.boundary_facets(T)
    return X, T, B, F, np.unique(F), np.flatnonzero(X[2] == 0)


@pytest.mark.parametrize("kind", KINDS)
def test_contact_oracle_supports_a_resting_body(kind):
    """Contact oracle sanity: a cube dropped on a fixed slab comes to rest on it."""
    X, T, B, F, V, dbc = _stacked()
    o = oracle.Oracle(X, T, dbc=dbc, B=B, V=V, F=F, muC=1e5, kind=kind)
    for _ in range(60):
        o.step(0.01, 10, 1)
    z = o.x[2, B == 1]
    assert z.min() > 0.45 and (o.get("nn") >= 0).any()




def test_xpbd_oracle_reference_doctest_and_header_build():
    """XPBD restatement (oracle/vbd_oracle.cpp, XpbdStep): the reference's known answer (sim/xpbd/Integrator.cpp:213-262: the
    unit cube falls, nothing moves sideways) and agreement of the port with the build that CALLS the reference's
    ProjectBlockNeoHookean / ProjectVertexTriangle (sim/xpbd/Kernels.h:78-141,169-246)."""
    from physicsbasedanimationtoolkit_b200 import meshes
    import physicsbasedanimationtoolkit_b200 as pbat

    P, T, F = meshes.CUBE_P, meshes.CUBE_T, meshes.CUBE_F
    kinds = ["port"] + (["reference"] if oracle.have_ref() else [])
    for kind in kinds:
        o = oracle.Oracle(P, T, V=np.arange(8), F=F, B=np.zeros(8, np.int64), kind=kind)
        o.xpbd_setup([0, 1, 2, 3, 4, 5], [0, 1, 2, 3, 4])
        o.xpbd_step(1e-2, 1, 20)
        dx = o.x - P
        assert (dx[2] < 0).all() and (np.abs(dx[:2]) < 1e-4).all()
    if len(kinds) == 2:
        X, T = meshes.tet_grid(8, 3, 3, 0.1)
        dbc = np.flatnonzero(X[0] == 0)
        Pptr, Padj, GC = pbat.sim.xpbd.partition_mesh_constraints(X, T)
        for c in range(GC.max() + 1):                      # a valid partitioning: no shared vertex inside a partition
            vs = T[:, GC == c].reshape(-1)
            assert vs.size == np.unique(vs).size
        out = []
        for kind in kinds:
            o = oracle.Oracle(X, T, dbc=dbc, kind=kind)
            o.xpbd_setup(Pptr, Padj, minv=np.full(X.shape[1], 0.1), beta_snh=np.full(2 * T.shape[1], 1e-3))
            for _ in range(5):
                o.xpbd_step(0.01, 5, 4)
            out.append(o.x)
        assert np.abs(out[0] - out[1]).max() < 1e-12 and np.abs(out[0] - X).max() > 1e-3
