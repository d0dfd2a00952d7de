"""``pbat.gpu.common.Buffer`` (bindings/pypbat/gpu/common/Buffer.cpp:24-61): a typed dims x n array handed to the
geometry / contact classes.  Here it is a host-side holder: the classes that consume it upload what they need (their
device state -- trees, boxes, active sets -- stays resident between calls)."""
from __future__ import annotations

import numpy as np


class Buffer:
    def __init__(self, data=None, n=None, dtype=np.float32, dims=1):
        if data is None:
            data = np.zeros((dims, 0 if n is None else n), dtype=dtype)
        elif np.isscalar(data):  # Buffer(dims, n, dtype)
            data = np.zeros((int(data), int(n)), dtype=dtype)
        self.set(data)

    def set(self, data):
        a = np.asarray(data)
        self._a = a.reshape(1, -1).copy() if a.ndim == 1 else a.copy()
        return self

    def resize(self, rows, cols=None):
        shape = (1, rows) if cols is None else (rows, cols)
        self._a = np.zeros(shape, dtype=self._a.dtype)

    def to_numpy(self):
        return self._a

    dims = property(lambda s: s._a.shape[0])
    size = property(lambda s: s._a.shape[1])
    type = property(lambda s: s._a.dtype)


def as_array(x, dtype, rows=None):
    """numpy view of a ``Buffer`` or array-like, ``rows`` x n."""
    a = x.to_numpy() if isinstance(x, Buffer) else np.asarray(x)
    a = np.asarray(a, dtype=dtype)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if rows is not None and a.shape[0] != rows:
        raise ValueError(f"expected an array with {rows} rows, got shape {a.shape}")
    return a
