"""GPU parity tests: the CUDA path (through the C-ABI / Python surface) against the CPU oracle.

Tolerance (BASELINE.json north_star): relative L2 position error <= 1e-4 after N steps, fp32
device arithmetic vs the double-precision reference semantics.  The stricter displacement-
relative error is reported alongside.
"""
import numpy as np
import pytest

import oracle
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

pytestmark = pytest.mark.gpu
TOL = 1e-4


def rel_l2(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def make(X, T, *, dbc=None, cheb=None, strategy=None, kD=0.0, omega_mode=0, tile_iters=0, klass=None, flags=0, v=None,
         kernel_variant=0, ring_slots=0, consumer_warps=0, material=0):
    d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_hyper_elastic_energy(material)
    if dbc is not None:
        d = d.with_dirichlet_vertices(dbc)
    if cheb:
        d = d.with_chebyshev_acceleration(cheb)
    if strategy is not None:
        d = d.with_initialization_strategy(strategy)
    if v is not None:
        d = d.with_velocity(v)
    d.omega_mode = omega_mode
    d = d.with_rayleigh_damping(kD).construct()
    klass = klass or pbat.gpu.vbd.Integrator
    vbd = klass(d, tile_iters=tile_iters, flags=flags, kernel_variant=kernel_variant, ring_slots=ring_slots,
                consumer_warps=consumer_warps)
    ref = oracle.Oracle(X, T, dbc=dbc, colors=d.colors, v=v,
                        accel=oracle.ACCEL_CHEBYSHEV if cheb else oracle.ACCEL_NONE, rho=cheb or 1.0,
                        omega_mode=omega_mode, kD=kD, material=material,
                        strategy=int(d.strategy))
    return d, vbd, ref


def test_cube_free_fall():
    """The reference's own doctest (sim/vbd/Integrator.cpp:245-293, gpu/impl/vbd/Integrator.cu:384-432)."""
    d, vbd, ref = make(meshes.CUBE_P, meshes.CUBE_T)
    vbd.step(1e-2, 10, 1)
    dx = vbd.x.astype(np.float64) - meshes.CUBE_P
    assert (dx[2] < 0).all()
    assert (np.abs(dx[:2]) < 1e-4).all()
    assert np.allclose(dx[2], -9.81e-4, atol=2e-6)
    ref.step(1e-2, 10, 1)
    assert rel_l2(vbd.x, ref.x) < TOL


def test_cube_chebyshev():
    """sim/vbd/ChebyshevIntegrator.cpp:40-85"""
    d, vbd, ref = make(meshes.CUBE_P, meshes.CUBE_T, cheb=0.9)
    vbd.step(1e-2, 10, 1)
    ref.step(1e-2, 10, 1)
    dx = vbd.x.astype(np.float64) - meshes.CUBE_P
    assert (dx[2] < 0).all() and (np.abs(dx[:2]) < 1e-4).all()
    assert rel_l2(vbd.x, ref.x) < TOL


def test_device_setup_matches_oracle():
    """vertex->tet CSR built on the device, GP / wg / m computed on the device."""
    X, T = meshes.tet_grid(7, 5, 4, 0.1)
    rng = np.random.default_rng(0)
    X = X + 0.01 * rng.uniform(-1, 1, X.shape)
    d, vbd, ref = make(X, T)
    p, e, il = vbd.adjacency()
    assert np.array_equal(p, ref.get("GVGp"))
    assert np.array_equal(e, ref.get("GVGe"))
    assert np.array_equal(il, ref.get("GVGilocal"))
    GP, wg, m = vbd.element_data()
    assert np.allclose(GP, ref.get("GP"), rtol=1e-12, atol=1e-12)
    assert np.allclose(wg, ref.get("wg"), rtol=1e-12)
    assert np.allclose(m, ref.get("m"), rtol=1e-12)
    assert np.array_equal(vbd.colors(), ref.get("colors"))


@pytest.mark.parametrize("cheb", [None, 0.9])
def test_config1_cantilever(cheb):
    """BASELINE.json configs[0]: Neo-Hookean cantilever, ~10k tets, 20 iters/step, 100 steps."""
    X, T = meshes.tet_grid(25, 9, 9, 0.04)
    dbc = np.flatnonzero(X[0] == 0)
    d, vbd, ref = make(X, T, dbc=dbc, cheb=cheb)
    worst = 0.0
    for s in range(100):
        vbd.step(0.01, 20, 1)
        ref.step(0.01, 20, 1)
        if s % 10 == 9:
            worst = max(worst, rel_l2(vbd.x, ref.x))
    xr = ref.x
    disp = np.linalg.norm(vbd.x - xr) / np.linalg.norm(xr - X)
    print(f"config1 cheb={cheb}: rel L2 = {rel_l2(vbd.x, xr):.3e} (worst {worst:.3e}), displacement-relative = {disp:.3e}, "
          f"tip deflection = {np.abs(xr - X).max():.3f} m")
    assert worst < TOL
    assert rel_l2(vbd.v, ref.v) < 1e-2  # velocities are differences of positions / dt


@pytest.mark.parametrize("strategy", list(pbat.sim.vbd.InitializationStrategy))
def test_initialization_strategies(strategy):
    X, T = meshes.tet_grid(6, 3, 3, 0.1)
    dbc = np.flatnonzero(X[0] == 0)
    v0 = np.zeros_like(X)
    v0[1] = 0.3 * X[0]
    d, vbd, ref = make(X, T, dbc=dbc, strategy=strategy, v=v0)
    for _ in range(10):
        vbd.step(0.01, 10, 1)
        ref.step(0.01, 10, 1)
    assert rel_l2(vbd.x, ref.x) < TOL


def test_substeps_damping_and_setters():
    X, T = meshes.tet_grid(6, 3, 3, 0.1)
    dbc = np.flatnonzero(X[0] == 0)
    d, vbd, ref = make(X, T, dbc=dbc, kD=1e-3, cheb=0.8, omega_mode=1)
    for _ in range(5):
        vbd.step(0.02, 8, 3)
        ref.step(0.02, 8, 3)
    assert rel_l2(vbd.x, ref.x) < TOL
    # state setters / getters round trip in the caller's vertex order
    x = vbd.x
    x[1] += 0.01
    vbd.x = x
    assert np.array_equal(vbd.x, x)
    ref.x = x.astype(np.float64)
    v = np.zeros_like(x)
    v[2] = 0.1
    v[:, dbc] = 0
    vbd.v = v
    ref.v = v.astype(np.float64)
    vbd.kD = 0.0
    ref.set_params(int(d.strategy), 0.0, 1e-7)
    vbd.step(0.01, 10, 1)
    ref.step(0.01, 10, 1)
    assert rel_l2(vbd.x, ref.x) < TOL


@pytest.mark.parametrize("tile_iters", [1, 2, 8])
def test_tile_shapes_agree(tile_iters):
    """The warp-tile shape is a pure scheduling choice: results must not depend on it beyond fp32
    summation order."""
    X, T = meshes.tet_grid(8, 4, 4, 0.1)
    dbc = np.flatnonzero(X[0] == 0)
    d, vbd, ref = make(X, T, dbc=dbc, tile_iters=tile_iters, cheb=0.9)
    for _ in range(10):
        vbd.step(0.01, 10, 1)
        ref.step(0.01, 10, 1)
    assert rel_l2(vbd.x, ref.x) < TOL


def test_sim_integrator_double_interface():
    X, T = meshes.tet_grid(5, 3, 3, 0.1)
    dbc = np.flatnonzero(X[0] == 0)
    d, vbd, ref = make(X, T, dbc=dbc, klass=pbat.sim.vbd.Integrator)
    vbd.step(0.01, 10)
    ref.step(0.01, 10, 1)
    assert vbd.x.dtype == np.float64 and rel_l2(vbd.x, ref.x) < TOL
    assert np.array_equal(vbd.data.x, vbd.x)


def test_errors():
    X, T = meshes.tet_grid(3, 2, 2, 0.1)
    d = pbat.sim.vbd.Data().with_volume_mesh(X, T).construct()
    bad = d.colors.copy()
    bad[:] = 0
    with pytest.raises(ValueError):
        pbat.gpu.vbd.Integrator(d, colors=bad)
    vbd = pbat.gpu.vbd.Integrator(d)
    with pytest.raises(ValueError):
        vbd.x = np.zeros((3, 5), dtype=np.float32)
    d2 = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_nesterov_acceleration(1.0, 3).construct()
    d2.nesterov_L = -1.0  # past Data::Construct's own check: the C-ABI validates again (sim/vbd/Data.cpp:284-293)
    with pytest.raises(ValueError):
        pbat.gpu.vbd.Integrator(d2)


@pytest.mark.parametrize("name", ["cube_base", "cube_cheb", "beam_small_base", "beam_small_cheb",
                                  "beam_small_cheb_textbook", "beam_small_substeps_damped", "beam_small_position",
                                  "beam_small_inertia", "beam_small_kinetic", "beam_small_adaptive_vbd",
                                  "beam_small_stvk", "beam_small_stvk_cheb_damped", "config1_base", "config1_cheb"])
def test_against_reference_golden(name):
    """CUDA path vs trajectories produced by the reference's own headers (tests/golden/make_golden.py)."""
    import os

    from golden.make_golden import CASES, mesh_of

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "vbd_golden.npz"))
    spec, kw, dt, iters, sub, steps = CASES[name]
    X, T, dbc = mesh_of(spec)
    d, vbd, _ = make(X, T, dbc=dbc, cheb=kw.get("rho") if kw.get("accel") else None, kD=kw.get("kD", 0.0),
                     omega_mode=kw.get("omega_mode", 0), material=kw.get("material", 0),
                     strategy=pbat.sim.vbd.InitializationStrategy(kw["strategy"]) if "strategy" in kw else None)
    assert np.array_equal(d.colors, gold[name + "/colors"])
    for _ in range(steps):
        vbd.step(dt, iters, sub)
    g = gold[name + "/x"]
    err = rel_l2(vbd.x, g)
    print(f"{name}: rel L2 vs reference golden = {err:.3e}")
    assert err < TOL


@pytest.mark.parametrize("cheb", [None, 0.9])
@pytest.mark.parametrize("ring_slots,consumer_warps", [(0, 0), (3, 0), (16, 16), (2, 3), (37, 7)])
def test_kernel_variants_bitwise_identical(cheb, ring_slots, consumer_warps):
    """The direct, TMA-ring, pipelined and one-cluster (4: the pipelined kernel behind the hardware cluster barrier)
    kernels do the same arithmetic in the same order: identical bits.  Small rings force many wrap-arounds of the
    producer/consumer pipeline."""
    X, T = meshes.tet_grid(12, 6, 5, 0.1)
    dbc = np.flatnonzero(X[0] == 0)
    out = []
    for variant in (1, 2, 3, 4):
        d, vbd, ref = make(X, T, dbc=dbc, cheb=cheb, kD=1e-4, kernel_variant=variant, ring_slots=ring_slots,
                           consumer_warps=consumer_warps, tile_iters=2)
        for _ in range(3):
            vbd.step(0.01, 7, 2)
        out.append((vbd.x, vbd.v))
    for o in out[1:]:
        assert np.array_equal(out[0][0], o[0]) and np.array_equal(out[0][1], o[1])
    for _ in range(3):
        ref.step(0.01, 7, 2)
    assert rel_l2(out[1][0], ref.x) < TOL


@pytest.mark.parametrize("cheb,material", [(None, 0), (0.9, 0), (0.8, 1), (None, 1)])
def test_barrier_free_sweep_is_bit_identical_to_the_barrier_sweep(cheb, material, monkeypatch):
    """The default whole-GPU sweep has no barrier between colours: positions carry the number of their write, tiles
    validate what they gathered and poll what is not there yet.  Same reads, same arithmetic: identical bits to the
    barrier sweep (VBDX_DATAFLOW=0), here with several tiles per warp and colour, substeps, and against the oracle."""
    X, T = meshes.tet_grid(34, 30, 26, 1 / 30)
    dbc = np.flatnonzero(X[2] == 0)
    x0 = X + 0.002 * np.random.default_rng(3).uniform(-1, 1, X.shape)
    x0[:, dbc] = X[:, dbc]
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("VBDX_DATAFLOW", mode)
        d, vbd, ref = make(X, T, dbc=dbc, cheb=cheb, kernel_variant=3, material=material)
        vbd.line_search_guard = material == 1 and cheb is None      # the guarded StVK step walks the records twice
        vbd.x = x0.astype(np.float32)
        for _ in range(4):
            vbd.step(0.01, 12, 2)
        out[mode] = (vbd.x.copy(), vbd.v.copy())
    assert np.array_equal(out["1"][0], out["0"][0]) and np.array_equal(out["1"][1], out["0"][1])
    ref.set_line_search_guard(material == 1 and cheb is None)
    ref.x = x0.astype(np.float32).astype(np.float64)
    for _ in range(4):
        ref.step(0.01, 12, 2)
    assert rel_l2(out["1"][0], ref.x) < TOL


@pytest.mark.parametrize("shape,kd", [((12, 10, 8), 0.0), ((12, 10, 8), 0.02), ((50, 44, 40), 0.02), ((70, 64, 60), 0.0)])
def test_three_sweep_kernels_agree_bitwise_over_sizes_and_damping(shape, kd, monkeypatch):
    """One scene, twenty steps, three schedules of the same arithmetic: the lean barrier-free kernel (default), the
    pipelined kernel sweeping barrier-free (VBDX_FLOW=0; it has no damping form, so it falls to barriers there) and the
    barrier sweep (VBDX_DATAFLOW=0).  From one tile per warp and colour up to a dozen; Rayleigh damping on and off."""
    X, T = meshes.tet_grid(*shape, 1 / shape[0])
    dbc = np.flatnonzero(X[2] == 0)
    x0 = X + 0.1 / shape[0] * np.random.default_rng(11).uniform(-1, 1, X.shape)
    x0[:, dbc] = X[:, dbc]
    out = []
    for env in ({}, {"VBDX_FLOW": "0"}, {"VBDX_DATAFLOW": "0"}):
        for k in ("VBDX_FLOW", "VBDX_DATAFLOW"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        d, vbd, _ = make(X, T, dbc=dbc, cheb=0.85, kernel_variant=3)
        vbd.kD = kd
        vbd.x = x0.astype(np.float32)
        for _ in range(20):
            vbd.step(0.01, 9, 1)
        out.append((vbd.x.copy(), vbd.v.copy(), vbd.info))
    assert out[0][2]["blockThreads"] != out[2][2]["blockThreads"]      # the lean kernel has no barrier warp
    for o in out[1:]:
        assert np.array_equal(out[0][0], o[0]) and np.array_equal(out[0][1], o[1])
    assert np.isfinite(out[0][0]).all()


def test_small_meshes_run_whole_steps_barrier_free_and_keep_the_cluster_for_barrier_launches(monkeypatch):
    """VBDX_KERNEL_DEFAULT on a mesh whose colours fit one thread-block cluster: whole steps are swept by the lean
    barrier-free kernel on the whole GPU (faster at every size since round 2); the 8-CTA cluster behind the hardware barrier
    serves what keeps colour barriers (VBDX_DATAFLOW=0, partial launches) and remains selectable (kernel_variant=4).
    Same bits whichever runs.  Contact, damping, substeps run on either."""
    X, T = meshes.tet_grid(25, 9, 9, 0.04)
    dbc = np.flatnonzero(X[0] == 0)
    d, vbd, ref = make(X, T, dbc=dbc, cheb=0.9, kD=1e-4)
    assert vbd.info["gridBlocks"] >= 100 and vbd.info["blockThreads"] % 32 == 0
    for _ in range(10):
        vbd.step(0.01, 20, 2)
        ref.step(0.01, 20, 2)
    assert rel_l2(vbd.x, ref.x) < TOL
    _, one_cluster, _ = make(X, T, dbc=dbc, cheb=0.9, kD=1e-4, kernel_variant=4)
    assert one_cluster.info["gridBlocks"] == 8
    monkeypatch.setenv("VBDX_DATAFLOW", "0")
    _, barriers, _ = make(X, T, dbc=dbc, cheb=0.9, kD=1e-4)
    assert barriers.info["gridBlocks"] == 8
    for other in (one_cluster, barriers):
        for _ in range(10):
            other.step(0.01, 20, 2)
        assert np.array_equal(other.x, vbd.x) and np.array_equal(other.v, vbd.v)
    monkeypatch.delenv("VBDX_DATAFLOW")
    Xl, Tl = meshes.tet_grid(40, 40, 40, 1 / 40)
    dl = pbat.sim.vbd.Data().with_volume_mesh(Xl, Tl).construct()
    assert pbat.gpu.vbd.Integrator(dl).info["gridBlocks"] >= 100


def test_direct_variant_config1():
    X, T = meshes.tet_grid(25, 9, 9, 0.04)
    dbc = np.flatnonzero(X[0] == 0)
    d, vbd, ref = make(X, T, dbc=dbc, cheb=0.9, kernel_variant=1)
    for s in range(20):
        vbd.step(0.01, 20, 1)
        ref.step(0.01, 20, 1)
    assert rel_l2(vbd.x, ref.x) < TOL


def test_host_layouts_and_pinned_buffers():
    """State crosses the boundary in either storage order and through page-locked arrays without change
    (vbdx_set/get_vertex_field; the f32/f64 column-major entry points stay the reference's convention)."""
    import ctypes as C
    from physicsbasedanimationtoolkit_b200 import _lib

    X, T = meshes.tet_grid(5, 4, 3, 0.1)
    nV = X.shape[1]
    d, vbd, _ = make(X, T, dbc=np.flatnonzero(X[2] == 0), cheb=0.7)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, nV)).astype(np.float32)
    vbd.x = x                                   # row-major
    assert np.array_equal(vbd.x, x)
    vbd.x = np.asfortranarray(x)                # column-major (Eigen), consumed in place
    assert np.array_equal(vbd.x, x)
    vbd.x = x[:, ::-1][:, ::-1]                 # non-contiguous view
    assert np.array_equal(vbd.x, x)
    pin = pbat.host.pinned_empty((3, nV), np.float32)
    pin[...] = x + 1
    vbd.x = pin
    out = pbat.host.pinned_empty((3, nV), np.float32)
    assert vbd.positions(out=out) is out and np.array_equal(out, x + 1)
    vbd.v = x
    assert np.array_equal(vbd.velocities(), x)
    # the legacy column-major C entry points agree with the generic ones
    L = _lib.lib()
    cols = np.empty((nV, 3), np.float32)
    _lib.check(L.vbdx_get_positions_f32(vbd._h, cols.ctypes.data, nV))
    assert np.array_equal(cols.T, x + 1)
    cols64 = np.empty((nV, 3), np.float64)
    _lib.check(L.vbdx_get_vertex_field(vbd._h, 0, 1, 0, cols64.ctypes.data, nV))
    assert np.array_equal(cols64.T, (x + 1).astype(np.float64))
    with pytest.raises(ValueError):
        vbd.positions(out=np.empty((nV, 3), np.float32))
    with pytest.raises(ValueError):
        _lib.check(L.vbdx_get_vertex_field(vbd._h, 2, 0, 0, cols.ctypes.data, nV))
    # double interface
    sim = pbat.sim.vbd.Integrator(d)
    xd = rng.standard_normal((3, nV))
    sim.x = xd
    assert np.array_equal(sim.x, xd.astype(np.float32).astype(np.float64))


def test_async_upload_step_download_equals_the_synchronous_calls():
    """vbdx_set_vertex_field_async / vbdx_step_async / vbdx_get_vertex_field_async / vbdx_synchronize: one host round trip per
    step instead of three (what bench.py's end-to-end leg does), same bits as the synchronous calls."""
    X, T = meshes.tet_grid(12, 10, 8, 0.1)
    dbc = np.flatnonzero(X[2] == 0)
    d, a, _ = make(X, T, dbc=dbc, cheb=0.9)
    _, b, _ = make(X, T, dbc=dbc, cheb=0.9)
    nV = X.shape[1]
    xin, xout = pbat.host.pinned_empty((3, nV), np.float32), pbat.host.pinned_empty((3, nV), np.float32)
    xin[...] = X + 0.003 * np.random.default_rng(2).uniform(-1, 1, X.shape)
    xin[:, dbc] = X[:, dbc]
    xs = xin.copy()
    for _ in range(6):
        a.set_positions_async(xin)
        a.step_async(0.01, 10, 1)
        a.positions_async(xout)
        a.synchronize()
        b.x = xs
        b.step(0.01, 10, 1)
        xs = b.x
        assert np.array_equal(xout, xs)
        xin, xout = xout, xin
    with pytest.raises(ValueError):
        a.set_positions_async(np.zeros((3, nV + 1), np.float32))


def test_anderson_acceleration():
    """AndersonIntegrator (sim/vbd/AndersonIntegrator.cpp:24-58): the reference's cube known answer and parity with
    the oracle on a cantilever where the window wraps around."""
    d = pbat.sim.vbd.Data().with_volume_mesh(meshes.CUBE_P, meshes.CUBE_T).with_anderson_acceleration(5).construct()
    vbd = pbat.sim.vbd.Integrator(d)
    dt = 1e-2
    x0 = vbd.x
    xtilde = x0 + dt * vbd.v + dt * dt * d.aext
    f0 = vbd.objective_function(x0, xtilde, dt)
    vbd.step(dt, 10, 1)
    dx = vbd.x - meshes.CUBE_P
    assert (dx[2] < 0).all() and (np.abs(dx[:2]) < 1e-4).all()
    assert np.linalg.norm(vbd.objective_function_gradient(vbd.x, xtilde, dt)) < 1e-4 * 50  # fp32 iterate, f in double
    assert vbd.objective_function(vbd.x, xtilde, dt) < f0

    X, T = meshes.tet_grid(12, 4, 4, 0.05)
    dbc = np.flatnonzero(X[0] == 0)
    for window, substeps in ((3, 1), (5, 2)):
        d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_anderson_acceleration(window).construct()
        vbd = pbat.gpu.vbd.Integrator(d)
        ref = oracle.Oracle(X, T, dbc=dbc, colors=d.colors)
        ref.set_acceleration(oracle.ACCEL_ANDERSON, window=window)
        plain = oracle.Oracle(X, T, dbc=dbc, colors=d.colors)
        # Tolerance: Anderson mixing amplifies the rounding of the iterates -- rounding the DOUBLE oracle's iterates
        # to fp32 after every sweep already moves this trajectory by 7e-5 after one step and 4.5e-4 after ten
        # (measured with the oracle alone).  Hence 2e-4 after the first step and 1.5e-3 after ten, instead of the
        # 1e-4 of the base and Chebyshev solves.
        for step in range(10):
            vbd.step(0.01, 10, substeps)
            ref.step(0.01, 10, substeps)
            plain.step(0.01, 10, substeps)
            if step == 0:
                assert rel_l2(vbd.x, ref.x) < 2e-4, rel_l2(vbd.x, ref.x)
        err = rel_l2(vbd.x, ref.x)
        assert err < 1.5e-3, err
        # and the acceleration really is in effect: the accelerated trajectory differs from the plain one by more
        assert rel_l2(plain.x, ref.x) > 3 * err
    with pytest.raises(ValueError):
        pbat.sim.vbd.Data().with_volume_mesh(X, T).with_anderson_acceleration(0).construct()


def test_nesterov_acceleration():
    """NesterovIntegrator::Solve (sim/vbd/NesterovIntegrator.cpp:18-44) against its literal restatement in the oracle
    (x^{k-1} captured once, sweep from x, correction x <- y^k - (x_swept - x^{k-1}) / L from iteration start + 1 on)."""
    X, T = meshes.tet_grid(12, 4, 4, 0.05)
    dbc = np.flatnonzero(X[0] == 0)
    # Restated literally, this iteration is not a contraction: the oracle's (double precision) displacement grows 20-fold
    # per step for L = 1 and still 7-fold for L = 4 on this cantilever (the reference's own Nesterov doctest fails for the
    # same reason, tests/test_oracle.py).  Parity is therefore checked step by step -- after every step the device state is
    # reset to the oracle's -- and relative to the step's UPDATE.  Tolerances: rounding the double oracle's iterates to
    # fp32 after every operation moves the update by 4.5e-3 / 8e-4 / 2e-4 (L = 4), 2.5e-2 / 1.4e-4 (L = 10, 2 substeps) and
    # 4.8e-3 / 2e-4 / 9e-6 (L = 1) in steps 0 / 1 / 2; the device differs from the oracle by the same amounts
    # (tools/nesterov_check.py prints both).  The first step moves the beam by 1e-4 of its length, hence the larger figure.
    for L, start, substeps, steps in ((4.0, 3, 1, 3), (10.0, 0, 2, 2), (1.0, 3, 1, 3)):
        d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_nesterov_acceleration(L, start).construct()
        vbd = pbat.sim.vbd.Integrator(d)
        ref = oracle.Oracle(X, T, dbc=dbc, colors=d.colors)
        ref.set_acceleration(oracle.ACCEL_NESTEROV, L=L, start=start)
        plain = oracle.Oracle(X, T, dbc=dbc, colors=d.colors)
        for step in range(steps):
            x0, v0 = ref.x, ref.v
            plain.x, plain.v = x0, v0
            vbd.x, vbd.v = x0, v0
            vbd.step(0.01, 10, substeps)
            ref.step(0.01, 10, substeps)
            plain.step(0.01, 10, substeps)
            assert np.isfinite(vbd.x).all()
            diff, upd = np.linalg.norm(vbd.x - ref.x), np.linalg.norm(ref.x - x0)
            assert diff < (6e-2 if step == 0 else 2e-3) * upd, (L, step, diff, upd)
            assert np.linalg.norm(plain.x - ref.x) > 5 * diff      # the accelerator is in effect
        assert rel_l2(vbd.x, ref.x) < 1e-4
    with pytest.raises(ValueError):
        pbat.sim.vbd.Data().with_volume_mesh(X, T).with_nesterov_acceleration(0.0, 3).construct()


def test_trust_region_acceleration():
    """TrustRegionIntegrator (gpu/impl/vbd/TrustRegionIntegrator.cu:35-262) against its restatement in the oracle.
    Curved path: the reference's constraint solver is a stub returning 0 (:700-703), no accelerated step can be taken and
    the iterates are the base solve's -- bit for bit here.  Linear path: parabola fit, clamped step, accept / reject."""
    X, T = meshes.tet_grid(10, 4, 4, 0.05)
    dbc = np.flatnonzero(X[0] == 0)
    base = pbat.gpu.vbd.Integrator(pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).construct())
    d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_trust_region_acceleration(0.2, 2.0, True).construct()
    curved = pbat.gpu.vbd.Integrator(d)
    for _ in range(3):
        base.step(0.01, 8, 1)
        curved.step(0.01, 8, 1)
    assert np.array_equal(base.x, curved.x)
    for eta, tau, substeps in ((0.2, 2.0, 1), (0.05, 1.5, 2)):
        d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_trust_region_acceleration(eta, tau, False).construct()
        vbd = pbat.gpu.vbd.Integrator(d)
        ref = oracle.Oracle(X, T, dbc=dbc, colors=d.colors)
        ref.set_trust_region(eta, tau, curved=False)
        plain = oracle.Oracle(X, T, dbc=dbc, colors=d.colors)
        # the accept / reject decisions compare differences of objective values; with fp32 iterates they stay the
        # oracle's as long as the solve is far from converged, hence few iterations per step
        for _ in range(4):
            vbd.step(0.01, 8, substeps)
            ref.step(0.01, 8, substeps)
            plain.step(0.01, 8, substeps)
        err = rel_l2(vbd.x, ref.x)
        assert np.isfinite(vbd.x).all() and err < 2e-4, err
        assert rel_l2(plain.x, ref.x) > 10 * err      # accelerated steps were taken
    with pytest.raises(ValueError):
        pbat.sim.vbd.Data().with_volume_mesh(X, T).with_trust_region_acceleration(0.2, 1.0, True).construct()


def test_broyden_acceleration():
    """BroydenIntegrator (sim/vbd/BroydenIntegrator.cpp:41-77): the reference's cube known answer (:83-135, 15
    iterations, m = 5) and parity with the oracle on a cantilever where the window wraps around."""
    d = pbat.sim.vbd.Data().with_volume_mesh(meshes.CUBE_P, meshes.CUBE_T).with_broyden_acceleration(5).construct()
    vbd = pbat.sim.vbd.Integrator(d)
    dt = 1e-2
    x0 = vbd.x
    xtilde = x0 + dt * vbd.v + dt * dt * d.aext
    f0 = vbd.objective_function(x0, xtilde, dt)
    vbd.step(dt, 15, 1)
    dx = vbd.x - meshes.CUBE_P
    assert (dx[2] < 0).all() and (np.abs(dx[:2]) < 1e-4).all()
    assert np.linalg.norm(vbd.objective_function_gradient(vbd.x, xtilde, dt)) < 1e-4 * 50  # fp32 iterate, f in double
    assert vbd.objective_function(vbd.x, xtilde, dt) < f0

    X, T = meshes.tet_grid(12, 4, 4, 0.05)
    dbc = np.flatnonzero(X[0] == 0)
    for window, substeps in ((3, 1), (5, 2)):
        d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_broyden_acceleration(window).construct()
        vbd = pbat.gpu.vbd.Integrator(d)
        ref = oracle.Oracle(X, T, dbc=dbc, colors=d.colors)
        ref.set_acceleration(oracle.ACCEL_BROYDEN, window=window)
        plain = oracle.Oracle(X, T, dbc=dbc, colors=d.colors)
        # Tolerance as for Anderson (same amplification of the fp32 rounding of the iterates by the secant mixing):
        # 2e-4 after the first step, 1.5e-3 after ten.
        for step in range(10):
            vbd.step(0.01, 10, substeps)
            ref.step(0.01, 10, substeps)
            plain.step(0.01, 10, substeps)
            if step == 0:
                print("broyden step 1:", rel_l2(vbd.x, ref.x))
                assert rel_l2(vbd.x, ref.x) < 2e-4, rel_l2(vbd.x, ref.x)
        err = rel_l2(vbd.x, ref.x)
        print("broyden step 10:", err, rel_l2(plain.x, ref.x))
        assert err < 1.5e-3, err
        assert rel_l2(plain.x, ref.x) > 3 * err
    with pytest.raises(ValueError):
        pbat.sim.vbd.Data().with_volume_mesh(X, T).with_broyden_acceleration(0).construct()


def test_objective_function_and_gradient():
    """Integrator::ObjectiveFunction / ObjectiveFunctionGradient (sim/vbd/Integrator.cpp:138-200) on the device
    (double precision) against the oracle, with per-element material parameters."""
    X, T = meshes.tet_grid(5, 3, 4, 0.2)
    nT = T.shape[1]
    rng = np.random.default_rng(5)
    mue, lame = 3e5 * (1 + rng.random(nT)), 2e6 * (1 + rng.random(nT))
    d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_material(np.full(nT, 1e3), mue, lame).construct()
    vbd = pbat.sim.vbd.Integrator(d)
    ref = oracle.Oracle(X, T, mue=mue, lambdae=lame, colors=d.colors)
    xk = X + 0.02 * rng.standard_normal(X.shape)
    xtilde = X + 0.01 * rng.standard_normal(X.shape)
    for dt in (1e-2, 0.3):
        f, g = vbd.objective_function(xk, xtilde, dt), vbd.objective_function_gradient(xk, xtilde, dt)
        fr, gr = ref.objective(xk, xtilde, dt), ref.objective_gradient(xk, xtilde, dt)
        assert abs(f - fr) <= 1e-12 * abs(fr)
        assert np.linalg.norm(g - gr.T.reshape(-1)) <= 1e-11 * np.linalg.norm(gr)
    d2 = pbat.sim.vbd.Data().with_volume_mesh(X, T).construct()          # default material path
    v2, r2 = pbat.sim.vbd.Integrator(d2), oracle.Oracle(X, T, colors=d2.colors)
    assert abs(v2.objective_function(xk, xtilde, 0.01) - r2.objective(xk, xtilde, 0.01)) <= 1e-12 * abs(r2.objective(xk, xtilde, 0.01))


@pytest.mark.parametrize("tile_iters", [0, 2])
def test_stvk_material(tile_iters):
    """St. Venant-Kirchhoff energy (physics/SaintVenantKirchhoffEnergy.h) in place of the Stable Neo-Hookean one:
    perturbed config-1-like beam with per-element Lame parameters, sweep + objective + gradient against the oracle."""
    X, T = meshes.tet_grid(12, 5, 4, 0.05)
    rng = np.random.default_rng(11)
    X = X + 0.004 * rng.uniform(-1, 1, X.shape)
    nT = T.shape[1]
    mue, lame = 3e5 * (1 + rng.random(nT)), 2e6 * (1 + rng.random(nT))
    dbc = np.flatnonzero(X[0] < 0.01)
    d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_material(np.full(nT, 1e3), mue, lame)
         .with_dirichlet_vertices(dbc).with_hyper_elastic_energy(pbat.sim.vbd.HyperElasticEnergy.SaintVenantKirchhoff)
         .with_chebyshev_acceleration(0.8).construct())
    vbd = pbat.sim.vbd.Integrator(d, tile_iters=tile_iters)
    ref = oracle.Oracle(X, T, mue=mue, lambdae=lame, dbc=dbc, colors=d.colors, accel=oracle.ACCEL_CHEBYSHEV, rho=0.8,
                        material=oracle.MATERIAL_STVK)
    for _ in range(20):
        vbd.step(0.01, 10, 1)
        ref.step(0.01, 10, 1)
    err = rel_l2(vbd.x, ref.x)
    derr = np.linalg.norm(vbd.x - ref.x) / np.linalg.norm(ref.x - X)
    print(f"stvk: rel L2 = {err:.3e}, displacement-relative = {derr:.3e}")
    assert err < TOL
    xk, xtilde = ref.x, ref.get("xtilde")
    f, g = vbd.objective_function(xk, xtilde, 0.01), vbd.objective_function_gradient(xk, xtilde, 0.01)
    fr, gr = ref.objective(xk, xtilde, 0.01), ref.objective_gradient(xk, xtilde, 0.01)
    assert abs(f - fr) <= 1e-12 * abs(fr)
    assert np.linalg.norm(g - gr.T.reshape(-1)) <= 1e-11 * np.linalg.norm(gr)
    with pytest.raises((RuntimeError, ValueError)):   # the direct / TMA variants carry Stable Neo-Hookean records only
        pbat.sim.vbd.Integrator(d, kernel_variant=1)


def test_line_search_guard():
    """The north star's "fused 3x3 Newton solve with line-search guard" (vbdx_set_line_search_guard; off by default =
    the reference's full Newton step).  (1) Stable Neo-Hookean: the local objective is exactly quadratic, the guard
    cannot alter a step: bit-identical.  (2) StVK near rest: every full step passes Armijo: bit-identical.  (3) StVK
    squashed to 60 % along x (indefinite Hessians): the unguarded sweep blows the objective up by orders of magnitude,
    the guarded one descends; checked against the oracle's restatement of the same guard."""
    X, T = meshes.tet_grid(8, 3, 3, 0.1)
    dbc = np.flatnonzero(X[0] == 0)
    E = pbat.sim.vbd.HyperElasticEnergy

    def build(energy, guard, x0=None):
        d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc)
             .with_hyper_elastic_energy(energy).construct())
        vbd = pbat.sim.vbd.Integrator(d)
        vbd.line_search_guard = guard
        if x0 is not None:
            vbd.x = x0
        return d, vbd

    rng = np.random.default_rng(4)
    xp = X + 0.002 * rng.uniform(-1, 1, X.shape)   # 2 % of a cell (at 10 % StVK is already unstable without the guard)
    xp[:, dbc] = X[:, dbc]
    for energy in (E.StableNeoHookean, E.SaintVenantKirchhoff):
        out = []
        for guard in (False, True):
            _, vbd = build(energy, guard, xp)
            for _ in range(5):
                vbd.step(0.01, 10, 1)
            out.append(vbd.x.copy())
        assert np.array_equal(out[0], out[1]), energy

    xs = X.copy()
    xs[0] *= 0.6
    xs[:, dbc] = X[:, dbc]
    f = {}
    for guard in (False, True):
        d, vbd = build(E.SaintVenantKirchhoff, guard, xs)
        ref = oracle.Oracle(X, T, dbc=dbc, colors=d.colors, material=oracle.MATERIAL_STVK)
        ref.set_line_search_guard(guard)
        ref.x = xs
        f[guard] = []
        for step in range(6):
            vbd.step(0.01, 10, 1)
            ref.step(0.01, 10, 1)
            xtilde = ref.get("xtilde")
            f[guard].append((vbd.objective_function(vbd.x, xtilde, 0.01), ref.objective(ref.x, xtilde, 0.01)))
        if guard:
            err = rel_l2(vbd.x, ref.x)
            print("guarded StVK, squashed: objective (gpu, oracle) per step:", f[guard], "rel L2 =", err)
            assert err < 1e-3
    fg = np.array(f[True])
    assert np.isfinite(fg).all() and (np.diff(fg[:, 0]) < 0).all() and fg[-1, 0] < 0.05
    assert np.allclose(fg[:, 0], fg[:, 1], rtol=0.05)
    assert max(v[1] for v in f[False]) > 100 * fg[0, 0]   # without the guard the same start diverges (oracle, double)


@pytest.mark.parametrize("cheb", [None, 0.8])
def test_traced_steps(tmp_path, cheb):
    """TraceNextStep / ExportTrace (sim/vbd/Integrator.cpp:47-52,202-235) and the GPU TracedStep
    (gpu/impl/vbd/Integrator.cu:105-148,284-301): a traced step is the same step, and the files hold the iterates."""
    from physicsbasedanimationtoolkit_b200 import mtx

    X, T = meshes.tet_grid(6, 3, 3, 0.1)
    dbc = np.flatnonzero(X[0] == 0)
    nV, iters, substeps = X.shape[1], 6, 2

    def data():
        d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc)
        return (d.with_chebyshev_acceleration(cheb) if cheb else d).construct()

    a, b = pbat.sim.vbd.Integrator(data()), pbat.sim.vbd.Integrator(data())
    a.step(0.01, iters, substeps), b.step(0.01, iters, substeps)
    a.trace_next_step(str(tmp_path), 7)
    a.step(0.01, iters, substeps)
    b.step(0.01, iters, substeps)
    assert np.array_equal(a.x, b.x) and np.array_equal(a.v, b.v)       # tracing does not change the step
    for s in range(substeps):
        f = mtx.load_dense(tmp_path / f"7.{s}.f.mtx")
        G = mtx.load_dense(tmp_path / f"7.{s}.grad.mtx")
        Xs = mtx.load_dense(tmp_path / f"7.{s}.x.mtx")
        assert f.shape == (iters + 1, 1) and G.shape == (3 * nV, iters + 1) and Xs.shape == (3 * nV, iters + 1)
        assert f[-1, 0] < f[0, 0]                                           # the solve descends
    assert np.array_equal(Xs[:, -1].reshape(nV, 3).T, a.x)
    a.step(0.01, iters, substeps)                                           # the flag was cleared
    assert not (tmp_path / "8.0.f.mtx").exists()
    # trajectory of the iterates against the oracle's plain sweeps (first substep, no acceleration)
    if not cheb:
        ref = oracle.Oracle(X, T, dbc=dbc, colors=a.data.colors)
        c = pbat.sim.vbd.Integrator(data())
        c.trace_next_step(str(tmp_path), 0)
        c.step(0.01, iters, 1)
        Xs = mtx.load_dense(tmp_path / "0.0.x.mtx")
        ref.step(0.01, 0, 1)
        ref.v = np.zeros_like(X)
        for k in range(iters):
            assert rel_l2(Xs[:, k].reshape(nV, 3).T, ref.x) < TOL
            ref.sweeps(0.01, 1)
    # GPU-flavoured trace
    g = pbat.gpu.vbd.Integrator(data())
    g.traced_step(0.01, iters, substeps, 0, str(tmp_path))
    h = pbat.gpu.vbd.Integrator(data())
    h.step(0.01, iters, substeps)
    assert np.array_equal(g.x, h.x)
    assert mtx.load_dense(tmp_path / "T.mtx").shape == (T.shape[1], 4)
    assert mtx.load_dense(tmp_path / "GP.mtx").shape == (4, 3 * T.shape[1])
    last = mtx.load_dense(tmp_path / f"x.t.0.s.{substeps - 1}.k.{iters}.mtx")
    assert last.shape == (nV, 3) and np.array_equal(last.T.astype(np.float32), g.x)
    assert mtx.load_dense(tmp_path / "xtilde.t.0.s.1.mtx").shape == (nV, 3)
