"""``pbat.gpu.contact.VertexTriangleMixedCcdDcd`` (bindings/pypbat/gpu/contact/VertexTriangleMixedCcdDcd.cpp:18-82): the
vertex-triangle detector of the VBD contact path on its own -- swept-box overlap over a linear BVH for the active set,
k nearest triangles of other bodies for the constraints (gpu/impl/contact/VertexTriangleMixedCcdDcd.cu:51-223)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .common import as_array

K_MAX_NEIGHBOURS = 8


class VertexTriangleMixedCcdDcd:
    def __init__(self, B, V, F):
        """``B``: body of every point, ``V``: collision vertices (subset of the points), ``F``: 3 x #triangles."""
        self._L = _lib.lib()
        self._h = C.c_void_p()
        B = np.ascontiguousarray(as_array(B, np.int64).reshape(-1))
        self._V = np.ascontiguousarray(as_array(V, np.int64).reshape(-1))
        Fa = np.ascontiguousarray(as_array(F, np.int64, 3).T)
        self.nV, self.nCV, self.nF = B.size, self._V.size, Fa.shape[0]
        _lib.check(self._L.vbdx_contact_create(self.nV, B.ctypes.data, self._V.ctypes.data, self.nCV, Fa.ctypes.data, self.nF,
                                               C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.vbdx_contact_destroy(self._h)
            self._h = None

    def _xyz(self, x):
        a = np.ascontiguousarray(as_array(x, np.float32, 3).T)
        if a.shape[0] != self.nV:
            raise ValueError(f"expected 3 x {self.nV} positions")
        return a

    def initialize_active_set(self, xt, xtp1, wmin, wmax):
        a, b = self._xyz(xt), self._xyz(xtp1)
        lo, hi = np.ascontiguousarray(wmin, np.float32).reshape(3), np.ascontiguousarray(wmax, np.float32).reshape(3)
        _lib.check(self._L.vbdx_contact_initialize_active_set(self._h, a.ctypes.data, b.ctypes.data, lo.ctypes.data, hi.ctypes.data))

    def update_active_set(self, x, bComputeBoxes=True):
        a = self._xyz(x)
        _lib.check(self._L.vbdx_contact_update_active_set(self._h, a.ctypes.data))

    def finalize_active_set(self, x, bComputeBoxes=True):
        a = self._xyz(x)
        _lib.check(self._L.vbdx_contact_finalize_active_set(self._h, a.ctypes.data))

    def _set_eps(self, eps):
        _lib.check(self._L.vbdx_contact_set_eps(self._h, float(eps)))

    eps = property(None, _set_eps, doc="floating point tolerance of the nearest neighbour search (write-only)")

    def _state(self):
        mask = np.zeros(self.nCV, np.int32)
        nn = np.zeros((self.nCV, K_MAX_NEIGHBOURS), np.int32)
        av = np.zeros(self.nCV, np.int32)
        na = C.c_int64(0)
        _lib.check(self._L.vbdx_contact_get(self._h, mask.ctypes.data, nn.ctypes.data, av.ctypes.data, C.byref(na)))
        return mask.astype(bool), nn, av[:na.value]

    @property
    def active_vertices(self):
        """Active vertex indices into ``V``."""
        return self._state()[2]

    @property
    def active_mask(self):
        return self._state()[0]

    @property
    def active_set(self):
        """2 x #constraints: (active vertex as an index into ``V``, triangle) pairs, as
        gpu/contact/VertexTriangleMixedCcdDcd.cu:93-121 assembles them."""
        _, nn, av = self._state()
        cols = []
        for v in av:
            for f in nn[v]:
                if f < 0:
                    break
                cols.append((v, f))
        return np.array(cols, dtype=np.int32).reshape(-1, 2).T
