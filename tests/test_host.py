"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol of
include/vbdx.h, host colouring equals the oracle's, the Python Data mirror reproduces
Data::Construct, and the library fails loudly (no fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import _lib, meshes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "vbdx.h")).read()
    declared = set(re.findall(r"\b(vbdx_[a-z0-9_A-Z]+)\s*\(", header))
    assert len(declared) >= 30
    L = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/vbdx.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert _lib.lib().vbdx_abi_version() == 2


def test_desc_struct_matches_header_size():
    d = _lib.DataDesc()
    _lib.lib().vbdx_data_desc_init(C.byref(d))
    assert d.struct_size == C.sizeof(_lib.DataDesc)
    assert (d.strategy, d.ordering, d.selection) == (4, 2, 0)
    assert (d.detHZero, d.muC, d.muF, d.epsv) == (1e-7, 1e6, 0.3, 1e-3)


@pytest.mark.parametrize("ordering", [0, 1, 2])
@pytest.mark.parametrize("selection", [0, 1])
def test_host_coloring_equals_oracle(ordering, selection):
    X, T = meshes.tet_grid(6, 5, 4)
    c = pbat.graph.mesh_greedy_color(T, X.shape[1], ordering, selection)
    o = oracle.Oracle(X, T, ordering=ordering, selection=selection)
    assert np.array_equal(c, o.get("colors"))


def test_coloring_irregular_mesh():
    rng = np.random.default_rng(1)
    X, T = meshes.tet_grid(5, 4, 3)
    keep = rng.uniform(size=T.shape[1]) > 0.3  # punch holes: irregular valences
    T = np.ascontiguousarray(T[:, keep])
    c = pbat.graph.mesh_greedy_color(T, X.shape[1])
    assert np.array_equal(c, oracle.Oracle(X, T).get("colors"))


def test_data_construct_matches_oracle():
    X, T = meshes.tet_grid(7, 4, 3, 0.1)
    X = X + 0.01 * np.random.default_rng(0).uniform(-1, 1, X.shape)
    dbc = np.flatnonzero(X[0] < 0.05)
    d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).construct()
    o = oracle.Oracle(X, T, dbc=dbc)
    for n in ("GVGp", "GVGe", "GVGilocal", "colors", "Pptr", "Padj"):
        assert np.array_equal(getattr(d, n), o.get(n)), n
    for n in ("m", "wg", "GP", "lame", "aext", "v"):
        assert np.allclose(getattr(d, n), o.get(n), rtol=1e-12, atol=1e-12), n
    assert d.strategy == pbat.sim.vbd.InitializationStrategy.AdaptivePbat


def test_data_surface_mesh_and_validation():
    d = pbat.sim.vbd.Data().with_volume_mesh(meshes.CUBE_P, meshes.CUBE_T).with_surface_mesh(np.arange(8), meshes.CUBE_F)
    assert np.isclose(d.FA.sum(), 6.0) and np.isclose(d.XVA.sum(), 6.0)
    with pytest.raises(ValueError):
        pbat.sim.vbd.Data().with_volume_mesh(meshes.CUBE_P, meshes.CUBE_T).with_chebyshev_acceleration(1.0).construct()
    with pytest.raises(ValueError):
        pbat.sim.vbd.Data().with_volume_mesh(meshes.CUBE_P, meshes.CUBE_T[[1, 0, 2, 3]]).construct()
    F = meshes.boundary_facets(meshes.CUBE_T)
    # outward orientation: normals point away from the centroid, as the reference's cube F does
    for FF in (F, meshes.CUBE_F):
        P = meshes.CUBE_P
        n = np.cross(P[:, FF[1]] - P[:, FF[0]], P[:, FF[2]] - P[:, FF[0]], axis=0)
        c = P[:, FF].mean(axis=1) - 0.5
        assert (np.einsum("ij,ij->j", n, c) > 0).all()


def test_mesh_generators():
    X, T = meshes.tet_grid(4, 3, 2, 0.25)
    assert T.shape[1] == 5 * 24 and (meshes.tet_volumes(X, T) > 0).all()
    assert np.isclose(meshes.tet_volumes(X, T).sum(), 24 * 0.25 ** 3)
    F = meshes.boundary_facets(T)
    assert F.shape[1] == 2 * 2 * (4 * 3 + 4 * 2 + 3 * 2)  # conforming: only the box surface is boundary
    Xs, Ts, B = meshes.stack_bodies(X, T, 3)
    assert Xs.shape[1] == 3 * X.shape[1] and B.max() == 2 and (meshes.tet_volumes(Xs, Ts) > 0).all()
    Xb, Tb = meshes.batch_scenes(X, T, 4, perturb=0.01)
    assert Tb.max() == 4 * X.shape[1] - 1 and (meshes.tet_volumes(Xb, Tb) > 0).all()


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of falling back."""
    if _lib.lib().vbdx_device_count() > 0:
        pytest.skip("a GPU is present")
    d = pbat.sim.vbd.Data().with_volume_mesh(meshes.CUBE_P, meshes.CUBE_T).construct()
    with pytest.raises(RuntimeError, match="no CUDA device"):
        pbat.gpu.vbd.Integrator(d)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "physicsbasedanimationtoolkit_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(base, f)).read()
                assert "oracle" not in src.replace("the parity oracle", "").replace("parity oracle", ""), f


def test_geometry_host_classes():
    """pbat.gpu.common.Buffer and pbat.gpu.geometry.Aabb are host-side holders: their logic needs no device."""
    Buffer, Aabb = pbat.gpu.common.Buffer, pbat.gpu.geometry.Aabb
    b = Buffer(np.arange(6, dtype=np.float32).reshape(3, 2))
    assert (b.dims, b.size, b.type) == (3, 2, np.float32)
    b.set(np.arange(4, dtype=np.int64))
    assert (b.dims, b.size) == (1, 4) and b.type == np.int64
    b.resize(3, 5)
    assert b.to_numpy().shape == (3, 5)
    a = Aabb(3, 0)
    a.construct(meshes.CUBE_P, meshes.CUBE_T)               # boxes of simplices (gpu/impl/geometry/Aabb.cu:18-74)
    assert a.n_boxes == 5 and a.dims == 3 and (a.min == 0).all() and (a.max == 1).all()
    P = np.array([[0., 2., 1.], [0., 0., 3.], [1., 1., 1.]])
    a.construct(Buffer(P.astype(np.float32)), Buffer(np.array([[0, 1], [1, 2]], dtype=np.int32)))   # two segments
    assert np.array_equal(a.min, [[0, 1], [0, 0], [1, 1]]) and np.array_equal(a.max, [[2, 2], [0, 3], [1, 1]])
    a.construct(np.zeros((3, 4)), np.ones((3, 4)))          # explicit min / max
    assert a.n_boxes == 4 and (a.max == 1).all()
    with pytest.raises(ValueError):
        a.construct(P, np.array([[0], [7]]))                # simplex index out of range
    with pytest.raises(ValueError):
        a.construct(np.zeros((3, 4)), np.ones((3, 5)))
    with pytest.raises(ValueError):
        Aabb(2, 4)


def test_batch_and_handles_fail_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d = pbat.sim.vbd.Data().with_volume_mesh(*meshes.tet_grid(2, 2, 2, 0.5)).construct()
    for make in (lambda: pbat.gpu.vbd.BatchIntegrator([d, d]), lambda: pbat.gpu.geometry.Bvh(8, 8),
                 lambda: pbat.gpu.contact.VertexTriangleMixedCcdDcd(np.zeros(4, int), np.arange(4), np.array([[0], [1], [2]]))):
        with pytest.raises(RuntimeError):
            make()
    with pytest.raises(ValueError):
        pbat.gpu.vbd.BatchIntegrator([])


def test_graph_helpers_and_xpbd_data():
    """pbat.graph.greedy_color on a compressed sparse graph equals the mesh colouring it generalises; map_to_adjacency;
    pbat.sim.xpbd.Data defaults and validation (sim/xpbd/Data.cpp:102-243)."""
    X, T = meshes.tet_grid(4, 3, 2, 0.1)
    G = pbat.graph.mesh_primal_graph(T, X.shape[1])
    for ordering in pbat.graph.GreedyColorOrderingStrategy:
        for selection in pbat.graph.GreedyColorSelectionStrategy:
            a = pbat.graph.greedy_color(G.indptr, G.indices, ordering, selection)
            b = pbat.graph.mesh_greedy_color(T, X.shape[1], ordering, selection)
            assert np.array_equal(a, b)
    ptr, adj = pbat.graph.map_to_adjacency(np.array([2, 0, 2, 1, 0]))
    assert ptr.tolist() == [0, 2, 3, 5] and adj.tolist() == [1, 4, 3, 0, 2]
    D = pbat.graph.mesh_dual_graph(T, X.shape[1])
    assert D.shape == (T.shape[1], T.shape[1]) and (D.diagonal() == 4).all()
    Pptr, Padj, GC = pbat.sim.xpbd.partition_mesh_constraints(X, T)
    d = pbat.sim.xpbd.Data().with_volume_mesh(X, T).with_partitions(Pptr, Padj).construct()
    assert np.allclose(d.minv, 1e-3) and (d.aext[2] == -9.81).all() and d.BV.sum() == 0 and d.gammaSNH.shape == (T.shape[1],)
    mu, lam = pbat.sim.vbd.lame_coefficients(1e6, 0.45)
    vol = np.abs(meshes.tet_volumes(X, T))
    assert np.allclose(d.alpha[0][0::2], 1 / (mu * vol)) and np.allclose(d.alpha[0][1::2], 1 / (lam * vol))
    e = 3
    Ds = np.stack([X[:, T[a, e]] - X[:, T[0, e]] for a in (1, 2, 3)], axis=1)
    assert np.allclose(d.DmInv[:, 3 * e:3 * e + 3] @ Ds, np.eye(3))
    with pytest.raises(ValueError):
        pbat.sim.xpbd.Data().with_volume_mesh(X, T).with_mass_inverse(np.ones(3)).construct()
    with pytest.raises(ValueError):
        pbat.sim.xpbd.Data().with_volume_mesh(X, T).with_surface_mesh(np.arange(4), np.zeros((3, 1), int)).with_collision_penalties(np.ones(2)).construct()


def test_device_colouring_refuses_the_sequential_selection_without_touching_a_gpu():
    """vbdx_greedy_color_device: the LeastUsed selection is inherently sequential (a global running count per colour) and is
    refused up front -- VBDX_UNSUPPORTED -> NotImplementedError, before any CUDA call (so also on a box without a GPU)."""
    X, T = meshes.tet_grid(3, 2, 2, 0.1)
    with pytest.raises(NotImplementedError, match="sequential"):
        pbat.graph.mesh_greedy_color(T, X.shape[1], 2, pbat.graph.GreedyColorSelectionStrategy.LeastUsed, device=0)
