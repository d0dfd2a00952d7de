"""``pbat.gpu.geometry``: ``Aabb`` and ``Bvh`` (bindings/pypbat/gpu/geometry/{Aabb,Bvh}.cpp) over the linear BVH of the
contact path (csrc/lbvh.cuh): Morton codes of the box centres, stable radix sort, Karras hierarchy with the reference's
node numbering, bottom-up boxes; self-overlap detection and nearest-triangle queries."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .common import as_array


class Aabb:
    """Axis-aligned boxes, ``dims`` x ``n_boxes`` (gpu/geometry/Aabb.h)."""

    def __init__(self, dims=3, n_aabb=0):
        self.resize(dims, n_aabb)

    def resize(self, dims, n_aabb):
        if dims != 3:
            raise ValueError("only 3-dimensional boxes are supported")
        self._lo = np.zeros((3, n_aabb), np.float32)
        self._hi = np.zeros((3, n_aabb), np.float32)

    def construct(self, a, b):
        """``construct(min, max)`` with two dims x n float arrays, or ``construct(P, S)`` with points P (3 x #pts) and
        simplices S (K x #simplices): the box of every simplex (gpu/impl/geometry/Aabb.cuh:58-78)."""
        b_arr = b.to_numpy() if hasattr(b, "to_numpy") else np.asarray(b)
        if np.issubdtype(b_arr.dtype, np.integer):
            P, S = as_array(a, np.float32, 3), as_array(b, np.int64)
            if S.size and (S.min() < 0 or S.max() >= P.shape[1]):
                raise ValueError("simplex index out of range")
            corners = P[:, S]                     # 3 x K x n
            self._lo, self._hi = corners.min(axis=1), corners.max(axis=1)
        else:
            L, U = as_array(a, np.float32, 3), as_array(b, np.float32, 3)
            if L.shape != U.shape:
                raise ValueError("min and max must have the same shape")
            self._lo, self._hi = L.copy(), U.copy()

    n_boxes = property(lambda s: s._lo.shape[1])
    dims = property(lambda s: 3)
    min = property(lambda s: s._lo)
    max = property(lambda s: s._hi)


class Bvh:
    """``Bvh(max_boxes, max_overlaps)`` (gpu/geometry/Bvh.h:33-148)."""

    def __init__(self, max_boxes, max_overlaps):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        self.max_boxes, self.max_overlaps = int(max_boxes), int(max_overlaps)
        _lib.check(self._L.vbdx_bvh_create(self.max_boxes, C.byref(self._h)))
        self._n = 0

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.vbdx_bvh_destroy(self._h)
            self._h = None

    def build(self, aabbs, min, max):
        L = np.ascontiguousarray(aabbs.min.T, np.float32)
        U = np.ascontiguousarray(aabbs.max.T, np.float32)
        wmin, wmax = np.ascontiguousarray(min, np.float32).reshape(3), np.ascontiguousarray(max, np.float32).reshape(3)
        _lib.check(self._L.vbdx_bvh_build(self._h, L.shape[0], L.ctypes.data, U.ctypes.data, wmin.ctypes.data, wmax.ctypes.data))
        self._n = L.shape[0]

    def detect_overlaps(self, aabbs, set=None):
        """Self-overlaps (bi < bj) of the boxes given to ``build`` as a 2 x #overlaps array; with ``set`` only pairs
        from different sets.  At most ``max_overlaps`` are returned."""
        if aabbs.n_boxes != self._n:
            raise ValueError("aabbs must be the one used in the last call to build()")
        sp = None
        if set is not None:
            sarr = np.ascontiguousarray(as_array(set, np.int32).reshape(-1))
            if sarr.size != self._n:
                raise ValueError("set must map every box to its set")
            sp = sarr.ctypes.data
        pairs = np.zeros((np.maximum(self.max_overlaps, 1), 2), np.int32)
        found = C.c_int64(0)
        _lib.check(self._L.vbdx_bvh_detect_overlaps(self._h, sp, self.max_overlaps, pairs.ctypes.data, C.byref(found)))
        self.n_overlaps_found = int(found.value)
        return pairs[:min_(self.n_overlaps_found, self.max_overlaps)].T.copy()

    def point_triangle_nearest_neighbours(self, aabbs, X, V, F):
        """Index of the nearest triangle of every column of ``X``; the tree must have been built over the boxes of the
        triangles ``F`` (3 x #triangles) of vertices ``V`` (3 x #verts)."""
        Xa = np.ascontiguousarray(as_array(X, np.float32, 3).T)
        Va = np.ascontiguousarray(as_array(V, np.float32, 3).T)
        Fa = np.ascontiguousarray(as_array(F, np.int32, 3).T)
        out = np.zeros(Xa.shape[0], np.int32)
        _lib.check(self._L.vbdx_bvh_nearest_triangles(self._h, Xa.shape[0], Xa.ctypes.data, Va.shape[0], Va.ctypes.data,
                                                      Fa.shape[0], Fa.ctypes.data, out.ctypes.data))
        return out

    def point_tetrahedron_nearest_neighbours(self, aabbs, X, V, T):
        raise NotImplementedError("nearest-tetrahedron queries are not on the VBD contact path")

    def _get(self):
        n = self._n
        if n < 1:
            raise RuntimeError("build() has not been called")
        ni = max(n - 1, 0)
        g = dict(child=np.zeros((2, ni), np.int32), parent=np.zeros(2 * n - 1, np.int32), rightmost=np.zeros((2, ni), np.int32),
                 inds=np.zeros(n, np.int32), codes=np.zeros(n, np.uint32), lo=np.zeros((2 * n - 1, 3), np.float32),
                 hi=np.zeros((2 * n - 1, 3), np.float32), visits=np.zeros(ni, np.int32))
        _lib.check(self._L.vbdx_bvh_get(self._h, *(g[k].ctypes.data for k in ("child", "parent", "rightmost", "inds", "codes", "lo", "hi", "visits"))))
        return g

    min = property(lambda s: s._get()["lo"].T.copy(), doc="BVH nodes' box minimums (3 x #nodes)")
    max = property(lambda s: s._get()["hi"].T.copy(), doc="BVH nodes' box maximums (3 x #nodes)")
    ordering = property(lambda s: s._get()["inds"], doc="box indices ordered by Morton code")
    morton = property(lambda s: s._get()["codes"], doc="sorted Morton codes")
    child = property(lambda s: s._get()["child"].T.copy(), doc="(#boxes - 1) x 2 children of every internal node")
    parent = property(lambda s: s._get()["parent"], doc="parents of all 2 #boxes - 1 nodes")
    rightmost = property(lambda s: s._get()["rightmost"].T.copy(), doc="(#boxes - 1) x 2 right-most leaves of the two subtrees")
    visits = property(lambda s: s._get()["visits"], doc="visits per internal node of the last box computation")


min_ = min


def build_bvh(L, U, wmin, wmax):
    """Build the LBVH over boxes ``L``/``U`` (3 x n) inside the world box.  Returns a dict with the
    reference's arrays: ``child`` (2 x (n-1)), ``parent`` (2n-1), ``rightmost`` (2 x (n-1)), ``inds`` (n),
    ``codes`` (n), ``lo``/``hi`` (3 x (2n-1) node boxes: internal nodes then leaves in sorted order)."""
    L = np.ascontiguousarray(np.asarray(L, np.float32).T)
    U = np.ascontiguousarray(np.asarray(U, np.float32).T)
    n = L.shape[0]
    wmin = np.ascontiguousarray(wmin, np.float32)
    wmax = np.ascontiguousarray(wmax, np.float32)
    child = np.zeros((2, max(n - 1, 0)), np.int32)
    right = np.zeros((2, max(n - 1, 0)), np.int32)
    parent = np.zeros(2 * n - 1, np.int32)
    inds = np.zeros(n, np.int32)
    codes = np.zeros(n, np.uint32)
    lo = np.zeros((2 * n - 1, 3), np.float32)
    hi = np.zeros((2 * n - 1, 3), np.float32)
    _lib.check(_lib.lib().vbdx_debug_bvh_build(n, L.ctypes.data, U.ctypes.data, wmin.ctypes.data, wmax.ctypes.data,
                                               child.ctypes.data, parent.ctypes.data, right.ctypes.data, inds.ctypes.data,
                                               codes.ctypes.data, lo.ctypes.data, hi.ctypes.data))
    return dict(child=child, parent=parent, rightmost=right, inds=inds, codes=codes, lo=lo.T.copy(), hi=hi.T.copy())
