"""Stand-alone access to the linear BVH the contact path uses (debug / parity surface modelled on
``pbat.gpu.geometry.Bvh``, bindings/pypbat/gpu/geometry/Bvh.cpp:19-109)."""
from __future__ import annotations

import numpy as np

from .. import _lib


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
