"""BASELINE.json configurations at their FULL sizes, through size-independent properties (the oracle needs about a
second per step at these sizes, so it is consulted for single steps only):

  configs[1]  58^3 cubes, 975,560 tets: analytic free fall, one full step against the oracle, run-to-run and
              kernel-variant bit-identity, monotone objective
  configs[2]  16 stacked bodies, 1.95 M tets, LBVH contact: the stack stays ordered, the first landing is caught by the penalty
  configs[4]  batch of independent 5k-tet scenes: a scene inside a batch evolves exactly like the scene alone
(configs[0] at full size: tests/test_gpu_parity.py; configs[3], domain decomposition: tests/test_gpu_dist.py.)
"""
import numpy as np
import pytest

import oracle
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

pytestmark = pytest.mark.gpu
GRID, ITERS, RHO, DT = 58, 30, 0.9, 0.01


def rel_l2(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.fixture(scope="module")
def config2():
    X, T = meshes.tet_grid(GRID, GRID, GRID, 1.0 / GRID)
    dbc = np.flatnonzero(X[2] == 0)
    x0 = X + 0.05 / GRID * np.random.default_rng(0).uniform(-1, 1, X.shape)
    x0[:, dbc] = X[:, dbc]
    d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc)
         .with_chebyshev_acceleration(RHO).construct())
    assert T.shape[1] == 975560 and X.shape[1] == 205379
    return X, T, dbc, x0, d


def test_config2_free_fall_known_answer():
    """The reference's cube doctest (sim/vbd/Integrator.cpp:245-293) at full size: an unconstrained body at rest falls
    rigidly, dz = -g dt^2 after the first step, nothing moves sideways.  Started from the inertial target
    (KineticEnergyMinimum) this holds at any size, because the elastic force of a rigid translate vanishes; from the
    default start it only holds once the sweeps have converged, which the 8-vertex cube does in 10 iterations and a
    10^6-tet block does not -- in the reference either."""
    X, T = meshes.tet_grid(GRID, GRID, GRID, 1.0 / GRID)
    d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_chebyshev_acceleration(RHO)
         .with_initialization_strategy(pbat.sim.vbd.InitializationStrategy.KineticEnergyMinimum).construct())
    vbd = pbat.gpu.vbd.Integrator(d)
    vbd.step(DT, 10, 1)
    dx = vbd.x.astype(np.float64) - X
    assert (dx[2] < 0).all()
    assert np.abs(dx[:2]).max() < 1e-4
    assert np.allclose(dx[2], -9.81e-4, atol=5e-6)
    assert np.allclose(vbd.v[2], -9.81e-2, atol=5e-4)


def test_config2_three_steps_against_the_oracle(config2):
    """BASELINE configs[1] at full size, three steps (90 sweeps of 198,651 vertices) against the double-precision oracle from
    the same fp32-representable start: position error and the stricter error relative to the DISPLACEMENT."""
    X, T, dbc, x0, d = config2
    vbd = pbat.gpu.vbd.Integrator(d)
    ref = oracle.Oracle(X, T, dbc=dbc, colors=d.colors, accel=oracle.ACCEL_CHEBYSHEV, rho=RHO)
    vbd.x = x0.astype(np.float32)
    ref.x = x0.astype(np.float32).astype(np.float64)
    for _ in range(3):
        vbd.step(DT, ITERS, 1)
        ref.step(DT, ITERS, 1)
    xr = ref.x
    err = rel_l2(vbd.x, xr)
    derr = np.linalg.norm(vbd.x - xr) / np.linalg.norm(xr - X)
    verr = np.linalg.norm(vbd.v - ref.v) / np.linalg.norm(ref.v)
    print(f"config 2, three full steps: rel L2 = {err:.3e}, displacement-relative = {derr:.3e}, velocity-relative = {verr:.3e}")
    assert err < 1e-4
    assert derr < 1e-3
    assert vbd.info["nonFiniteVertices"] == 0


def test_config2_bit_identical_across_runs_and_kernel_variants(config2):
    X, T, dbc, x0, d = config2
    out = []
    for variant in (0, 0, 1):  # default (pipelined) twice, then the direct kernel
        vbd = pbat.gpu.vbd.Integrator(d, kernel_variant=variant)
        vbd.x = x0.astype(np.float32)
        for _ in range(3):
            vbd.step(DT, ITERS, 1)
        out.append((vbd.x.copy(), vbd.v.copy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][0], out[2][0]) and np.array_equal(out[0][1], out[2][1])


def test_config2_sweeps_descend(config2):
    """Block coordinate Newton steps on the backward-Euler objective: every batch of sweeps lowers f and the
    gradient norm, at full size (objective evaluated on the device in double)."""
    X, T, dbc, x0, _ = config2
    d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).construct()
    vbd = pbat.sim.vbd.Integrator(d)
    vbd.x = x0
    free = np.ones(X.shape[1], bool)
    free[dbc] = False
    xtilde = x0 + DT * DT * d.aext
    f = [vbd.objective_function(x0, xtilde, DT)]
    g = [np.linalg.norm(vbd.objective_function_gradient(x0, xtilde, DT).reshape(-1, 3)[free])]
    for k0 in range(0, 12, 4):
        vbd.step_partial(DT, k0, k0 + 4, 12, flags=(1 if k0 == 0 else 0))
        xk = vbd.x
        f.append(vbd.objective_function(xk, xtilde, DT))
        g.append(np.linalg.norm(vbd.objective_function_gradient(xk, xtilde, DT).reshape(-1, 3)[free]))
    print("f:", f, "|grad|:", g)
    assert all(b < a for a, b in zip(f, f[1:]))
    assert g[-1] < 0.05 * g[0]


def test_config3_stack_stays_ordered_and_lands():
    n = 29
    Xb, Tb = meshes.tet_grid(n, n, n, 1.0 / n)
    X, T, B = meshes.stack_bodies(Xb, Tb, 16, axis=2, gap_frac=0.1)
    assert T.shape[1] == 1951120
    F = meshes.boundary_facets(T)
    V = np.unique(F)
    assert F.shape[1] == 161472 and V.size == 80768
    dbc = np.flatnonzero(X[2] <= X[2].min() + 0.01)
    d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_surface_mesh(V, F).with_bodies(B)
         .with_dirichlet_vertices(dbc).with_contact_parameters(1e6, 0.3, 1e-3).construct())
    vbd = pbat.gpu.vbd.Integrator(d)
    for _ in range(40):
        vbd.step(DT, 20, 1)
    x = vbd.x
    assert np.isfinite(x).all()
    lo = np.array([x[2, B == b].min() for b in range(16)])
    hi = np.array([x[2, B == b].max() for b in range(16)])
    assert (np.diff(lo) > 0).all()
    # Penalty contact (muC = 1e6): a surface vertex carries k = muC * area ~ 1.2e3 N/m against the ~1.2 kg column
    # above it, i.e. a static depth of ~0.01 = 0.3 cells per resting body and twice that on impact.  After 40 steps
    # only body 1 has landed (on the fixed body 0): it may be in by less than a cell; every other gap is still open.
    gaps = lo[1:] - hi[:-1]
    assert gaps[0] > -1.0 / n, gaps[0]
    assert (gaps[1:] > 0).all(), gaps
    assert gaps[0] < 0.01  # ... and it HAS landed (free fall alone would have carried it 0.7 below body 0's top)
    active, nn, n_active = vbd.contact_state()
    # nActive = size of the step's compacted candidate list (InitializeActiveSet); FinalizeActiveSet then clears the
    # flags of the vertices that ended on the positive side of their nearest triangle
    assert 0 < np.count_nonzero(active) <= n_active < V.size // 10
    tri = nn[nn >= 0]
    assert (tri < F.shape[1]).all()
    # a contact pairs a vertex with a triangle of ANOTHER body (VertexTriangleMixedCcdDcd.cuh:108-162)
    vi, slot = np.nonzero(nn >= 0)
    assert (B[V[vi]] != B[F[0, nn[vi, slot]]]).all()


def test_config3_barrier_free_equals_colour_barriers_at_full_size(monkeypatch):
    """configs[2] at its size, 45 steps (free fall, first impacts, ~2,000 active vertices): the barrier-free sweep with the
    write history (default) and the sweep with colour barriers take the same contact decisions and produce the same bits
    -- also on the degenerate geometry of identical grids stacked exactly above each other, where one ulp flips a contact."""
    n = 29
    Xb, Tb = meshes.tet_grid(n, n, n, 1.0 / n)
    X, T, B = meshes.stack_bodies(Xb, Tb, 16, axis=2, gap_frac=0.1)
    F = meshes.boundary_facets(T)
    V = np.unique(F)
    dbc = np.flatnonzero(X[2] <= X[2].min() + 0.01)
    d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_surface_mesh(V, F).with_bodies(B)
         .with_dirichlet_vertices(dbc).with_contact_parameters(1e6, 0.3, 1e-3).construct())
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("VBDX_DATAFLOW", mode)
        vbd = pbat.gpu.vbd.Integrator(d)
        for _ in range(45):
            vbd.step(DT, 20, 1)
        out[mode] = (vbd.x.copy(), vbd.v.copy(), vbd.contact_state(), vbd.info["blockThreads"])
    assert out["1"][3] != out["0"][3]                       # the lean kernel (no barrier warp) ran the default
    assert out["1"][2][2] > 1000                            # vertices in the active set
    assert np.array_equal(out["1"][0], out["0"][0]) and np.array_equal(out["1"][1], out["0"][1])
    assert np.array_equal(out["1"][2][0], out["0"][2][0]) and np.array_equal(out["1"][2][1], out["0"][2][1])


def test_config5_scene_in_a_batch_equals_the_scene_alone():
    n_scenes = 256
    Xs, Ts = meshes.tet_grid(10, 10, 10, 0.1)
    assert Ts.shape[1] == 5000
    X, T = meshes.batch_scenes(Xs, Ts, n_scenes, perturb=0.002)
    nV = Xs.shape[1]
    dbc1 = np.flatnonzero(Xs[2] == 0)
    dbc = (dbc1[None, :] + nV * np.arange(n_scenes)[:, None]).reshape(-1)
    single = pbat.sim.vbd.Data().with_volume_mesh(Xs, Ts).with_dirichlet_vertices(dbc1).construct()
    colors = np.tile(single.colors, n_scenes)
    d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).construct()
    batch = pbat.gpu.vbd.Integrator(d, colors=colors)
    for _ in range(10):
        batch.step(DT, 20, 1)
    xb = batch.x
    assert np.isfinite(xb).all()
    for s in (0, 101, n_scenes - 1):
        Xi = X[:, s * nV:(s + 1) * nV]
        di = pbat.sim.vbd.Data().with_volume_mesh(Xi, Ts).with_dirichlet_vertices(dbc1).construct()
        alone = pbat.gpu.vbd.Integrator(di, colors=single.colors)
        for _ in range(10):
            alone.step(DT, 20, 1)
        assert np.array_equal(alone.x, xb[:, s * nV:(s + 1) * nV]), s


def test_batch_entry_point_with_ragged_scenes():
    """vbdx_create_batch (pbat.gpu.vbd.BatchIntegrator): scenes of different meshes, materials and constraints in one
    batch; each evolves bit-identically to the scene stepped alone through vbdx_create."""
    rng = np.random.default_rng(7)
    datas = []
    for s in range(48):
        nx, ny, nz = ((10, 10, 10), (6, 4, 3), (12, 3, 2), (2, 2, 2))[s % 4]
        X, T = meshes.tet_grid(nx, ny, nz, 0.1)
        X = X + 0.002 * rng.uniform(-1, 1, X.shape)
        d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_chebyshev_acceleration(0.8)
        if s % 3:
            d = d.with_dirichlet_vertices(np.flatnonzero(X[0] < 0.05))
        if s % 5 == 0:
            nT = T.shape[1]
            d = d.with_material(np.full(nT, 800.0 + 10 * s), np.full(nT, 2e5 * (1 + s % 4)), np.full(nT, 2e6))
        if s % 7 == 0:
            d = d.with_velocity(0.1 * rng.standard_normal(X.shape))
        datas.append(d.construct())
    batch = pbat.gpu.vbd.BatchIntegrator(datas)
    assert batch.n_scenes == 48 and batch.offsets[-1] == sum(d.X.shape[1] for d in datas)
    for _ in range(8):
        batch.step(DT, 15, 2)
    xb, vb = batch.x, batch.v
    assert np.isfinite(xb).all()
    for s in (0, 1, 2, 3, 5, 14, 35, 47):
        alone = pbat.gpu.vbd.Integrator(datas[s])
        for _ in range(8):
            alone.step(DT, 15, 2)
        assert np.array_equal(alone.x, batch.scene(xb, s)), s
        assert np.array_equal(alone.v, batch.scene(vb, s)), s
    # settings must agree across the scenes; contact is per-integrator only
    other = pbat.sim.vbd.Data().with_volume_mesh(*meshes.tet_grid(2, 2, 2, 0.1)).construct()
    with pytest.raises(ValueError):
        pbat.gpu.vbd.BatchIntegrator([datas[0], other])


def test_scene_batches_sharded_over_devices():
    """BASELINE configs[4] at reduced size: scenes block-partitioned over several handles / devices (no communication) evolve
    bit-identically to the same scenes in ONE batch.  Uses every visible GPU, or two handles on the only one."""
    from physicsbasedanimationtoolkit_b200 import _lib

    Xs, Ts = meshes.tet_grid(6, 6, 6, 0.1)
    fixed = np.flatnonzero(Xs[2] == 0)
    datas = []
    for s in range(24):
        xs = Xs + 0.002 * np.random.default_rng(s).uniform(-1, 1, Xs.shape)
        datas.append(pbat.sim.vbd.Data().with_volume_mesh(xs, Ts).with_dirichlet_vertices(fixed).construct())
    n_dev = _lib.lib().vbdx_device_count()
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]
    one = pbat.gpu.vbd.BatchIntegrator(datas)
    many = pbat.gpu.vbd.MultiGpuBatchIntegrator(datas, devices=devices)
    assert many.n_scenes == 24 and many.nV == one.nV and np.array_equal(many.offsets, one.offsets)
    for _ in range(5):
        one.step(0.01, 10, 1)
        many.step(0.01, 10, 1)
    assert np.array_equal(one.x, many.x) and np.array_equal(one.v, many.v)
    x = many.x
    x[2] += 0.01
    many.x = x
    assert np.array_equal(many.x, x.astype(np.float32))
    assert np.array_equal(many.scene(x, 5), x[:, one.offsets[5]:one.offsets[6]])
