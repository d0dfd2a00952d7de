"""The stand-alone geometry / contact surfaces (pbat.gpu.geometry.{Aabb,Bvh}, pbat.gpu.contact.VertexTriangleMixedCcdDcd,
pbat.gpu.common.Buffer): the reference's own known answers (gpu/impl/geometry/Bvh.cu:337-580,
gpu/impl/geometry/Aabb.cu:18-74, gpu/impl/contact/VertexTriangleMixedCcdDcd.cu:232-329) and brute-force cross-checks."""
import numpy as np
import pytest

import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

pytestmark = pytest.mark.gpu


def test_aabb_and_bvh_golden_topology():
    """The 5 tets of the unit cube all have the unit box: duplicate Morton codes, tie-break by leaf index."""
    P, T = meshes.CUBE_P, meshes.CUBE_T
    aabbs = pbat.gpu.geometry.Aabb(3, T.shape[1])
    aabbs.construct(P, T)
    assert aabbs.n_boxes == 5 and (aabbs.min == 0).all() and (aabbs.max == 1).all()      # Aabb.cu:18-74
    bvh = pbat.gpu.geometry.Bvh(T.shape[1], 10)
    bvh.build(aabbs, P.min(axis=1), P.max(axis=1))
    assert bvh.child.tolist() == [[3, 8], [4, 5], [6, 7], [1, 2]]                       # Bvh.cu:367-410
    assert bvh.parent.tolist() == [-1, 3, 3, 0, 1, 1, 2, 2, 0]
    assert bvh.rightmost.tolist() == [[7, 8], [4, 5], [6, 7], [5, 7]]
    assert bvh.ordering.tolist() == [0, 1, 2, 3, 4] and (bvh.visits == 2).all()
    assert (bvh.min == 0).all() and (bvh.max == 1).all()
    # every pair of the 5 identical boxes overlaps: 10 pairs, bi < bj, each once
    O = bvh.detect_overlaps(aabbs)
    assert O.shape == (2, 10) and (O[0] < O[1]).all()
    assert len({tuple(c) for c in O.T}) == 10


@pytest.mark.parametrize("n", [2, 3, 100, 5000])
def test_detect_overlaps_equals_brute_force(n):
    rng = np.random.default_rng(n)
    c = rng.uniform(0, 1, (3, n)).astype(np.float32)
    h = rng.uniform(0, 0.5 / n ** (1 / 3), (3, n)).astype(np.float32)
    aabbs = pbat.gpu.geometry.Aabb()
    aabbs.construct(c - h, c + h)
    L, U = aabbs.min, aabbs.max
    bvh = pbat.gpu.geometry.Bvh(n, 40 * n)
    bvh.build(aabbs, [0, 0, 0], [1, 1, 1])
    ov = (L[:, :, None] <= U[:, None, :]).all(axis=0) & (U[:, :, None] >= L[:, None, :]).all(axis=0)
    i, j = np.nonzero(np.triu(ov, 1))
    expected = set(zip(i.tolist(), j.tolist()))
    O = bvh.detect_overlaps(aabbs)
    got = [tuple(c) for c in O.T.tolist()]
    assert len(got) == len(set(got)) and set(got) == expected
    sets = rng.integers(0, 3, n).astype(np.int32)
    O2 = bvh.detect_overlaps(aabbs, pbat.gpu.common.Buffer(sets))
    assert {tuple(c) for c in O2.T.tolist()} == {(a, b) for a, b in expected if sets[a] != sets[b]}
    if len(expected) > 3:                      # a too small output buffer: truncated, the true count is still reported
        small = pbat.gpu.geometry.Bvh(n, 3)
        small.build(aabbs, [0, 0, 0], [1, 1, 1])
        O3 = small.detect_overlaps(aabbs)
        assert O3.shape == (2, 3) and small.n_overlaps_found == len(expected)
        assert {tuple(c) for c in O3.T.tolist()} <= expected


def point_triangle_d2(p, A, B, C):
    """Brute force squared distance by dense barycentric sampling refined with the exact edge/vertex cases."""
    def seg(p, a, b):
        ab = b - a
        t = np.clip(((p - a) * ab).sum(-1) / (ab * ab).sum(-1), 0, 1)
        q = a + t[..., None] * ab
        return ((p - q) ** 2).sum(-1)
    n = np.cross(B - A, C - A)
    nn = (n * n).sum(-1)
    d = ((p - A) * n).sum(-1) / nn
    q = p - d[..., None] * n
    def inside(a, b):
        return (np.cross(b - a, q - a) * n).sum(-1) >= 0
    ins = inside(A, B) & inside(B, C) & inside(C, A)
    plane = d * d * nn
    edges = np.minimum(np.minimum(seg(p, A, B), seg(p, B, C)), seg(p, C, A))
    return np.where(ins, plane, edges)


def test_point_triangle_nearest_neighbours():
    X0, T = meshes.tet_grid(6, 5, 4, 0.25)
    F = meshes.boundary_facets(T)
    V = X0 + 0.03 * np.random.default_rng(0).uniform(-1, 1, X0.shape)
    aabbs = pbat.gpu.geometry.Aabb(3, F.shape[1])
    aabbs.construct(V, F)
    bvh = pbat.gpu.geometry.Bvh(F.shape[1], 0)
    bvh.build(aabbs, V.min(axis=1) - 1, V.max(axis=1) + 1)
    Q = np.random.default_rng(1).uniform(-0.5, 2.0, (3, 400))
    nn = bvh.point_triangle_nearest_neighbours(aabbs, pbat.gpu.common.Buffer(Q.astype(np.float32)), pbat.gpu.common.Buffer(V.astype(np.float32)),
                                               pbat.gpu.common.Buffer(F.astype(np.int32)))
    assert nn.shape == (400,) and (nn >= 0).all() and (nn < F.shape[1]).all()
    Vf, Qf = V.astype(np.float32).astype(np.float64), Q.astype(np.float32).astype(np.float64)
    A, B, C = (Vf[:, F[k]].T for k in range(3))
    for q in range(Q.shape[1]):
        d2 = point_triangle_d2(Qf[:, q][None, :], A, B, C)
        assert d2[nn[q]] <= d2.min() * (1 + 1e-4) + 1e-9, (q, d2[nn[q]], d2.min())


def test_vertex_triangle_detector_two_tets():
    """gpu/impl/contact/VertexTriangleMixedCcdDcd.cu:232-329."""
    XT = np.array([[0., 1., 0., 0.1, 0., 1., 0., 0.1],
                   [0., 0., 1., 0.1, 0., 0., 1., 0.1],
                   [0., 0., 0., 1., 1.01, 1.01, 1.01, 2.01]], np.float32)
    F = np.array([[0, 1, 2, 0, 4, 5, 6, 4], [1, 2, 0, 2, 5, 6, 4, 6], [3, 3, 3, 1, 7, 7, 7, 5]], dtype=np.int64)
    V = np.arange(8)
    B = np.array([0, 0, 0, 0, 1, 1, 1, 1])
    X = XT.copy()
    X[2, :4] += 0.01
    X[2, 4:] -= 0.01
    ccd = pbat.gpu.contact.VertexTriangleMixedCcdDcd(B, V, F)
    ccd.initialize_active_set(pbat.gpu.common.Buffer(XT), pbat.gpu.common.Buffer(X), [0, 0, 0], [1, 1, 2.01])
    mask, av = ccd.active_mask, ccd.active_vertices
    assert av.size == 4 and mask.sum() == 4 and (av == 3).any()
    ccd.update_active_set(pbat.gpu.common.Buffer(X))
    A = ccd.active_set
    assert A.shape[0] == 2 and A.shape[1] >= 4
    with_nn = set(A[0].tolist())
    assert with_nn == set(np.flatnonzero(mask).tolist())                 # active <=> has nearest neighbours
    assert (B[V[A[0]]] != B[F[0, A[1]]]).all()                           # always a triangle of the other body
    ccd.eps = 1e-6
    ccd.finalize_active_set(X)
    assert ccd.active_mask.sum() <= 4
