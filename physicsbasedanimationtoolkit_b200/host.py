"""Page-locked host arrays for the integrators' ``x``/``v`` traffic (``vbdx_host_alloc``): a pinned
array handed to ``Integrator.x = ...`` or ``Integrator.positions(out=...)`` is transferred by one DMA copy
without the driver's pageable staging."""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib


def pinned_empty(shape, dtype=np.float32):
    """Uninitialised C-contiguous numpy array in page-locked memory; freed with the array."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    nbytes = max(n * dtype.itemsize, 1)
    L = _lib.lib()
    p = C.c_void_p()
    _lib.check(L.vbdx_host_alloc(C.byref(p), nbytes))
    buf = (C.c_char * nbytes).from_address(p.value)
    a = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    weakref.finalize(buf, L.vbdx_host_free, p.value)
    return a
