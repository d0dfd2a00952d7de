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
         kernel_variant=0, ring_slots=0, consumer_warps=0):
    d = pbat.sim.vbd.Data().with_volume_mesh(X, T)
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
                        omega_mode=omega_mode, kD=kD,
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
    d2 = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_anderson_acceleration(5).construct()
    with pytest.raises(NotImplementedError):
        pbat.gpu.vbd.Integrator(d2)


@pytest.mark.parametrize("name", ["cube_base", "cube_cheb", "beam_small_base", "beam_small_cheb",
                                  "beam_small_cheb_textbook", "beam_small_substeps_damped", "beam_small_position",
                                  "beam_small_inertia", "beam_small_kinetic", "beam_small_adaptive_vbd",
                                  "config1_base", "config1_cheb"])
def test_against_reference_golden(name):
    """CUDA path vs trajectories produced by the reference's own headers (tests/golden/make_golden.py)."""
    import os

    from golden.make_golden import CASES, mesh_of

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "vbd_golden.npz"))
    spec, kw, dt, iters, sub, steps = CASES[name]
    X, T, dbc = mesh_of(spec)
    d, vbd, _ = make(X, T, dbc=dbc, cheb=kw.get("rho") if kw.get("accel") else None, kD=kw.get("kD", 0.0),
                     omega_mode=kw.get("omega_mode", 0),
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
    """The direct, TMA-ring and pipelined kernels do the same arithmetic in the same order: identical bits.
    Small rings force many wrap-arounds of the producer/consumer pipeline."""
    X, T = meshes.tet_grid(12, 6, 5, 0.1)
    dbc = np.flatnonzero(X[0] == 0)
    out = []
    for variant in (1, 2, 3):
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
