"""Dense Matrix Market files as ``Eigen::saveMarketDense`` writes them (unsupported/Eigen/src/SparseExtra/MarketIO.h):
``%%MatrixMarket matrix array real general``, ``rows cols``, then the coefficients in column-major order, one per
line.  The reference's iterate traces use this format (sim/vbd/Integrator.cpp:202-235,
gpu/impl/vbd/Integrator.cu:105-148,284-301; consumers python/vbd/convergence.py, python/vbd/path.py)."""
from __future__ import annotations

import numpy as np


def save_dense(path, a):
    a = np.atleast_2d(np.asarray(a))
    if a.ndim != 2:
        raise ValueError("expected a vector or a matrix")
    integer = np.issubdtype(a.dtype, np.integer)
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix array {'integer' if integer else 'real'} general\n{a.shape[0]} {a.shape[1]}\n")
        np.savetxt(f, a.T.reshape(-1), fmt="%d" if integer else "%.17g")


def load_dense(path):
    with open(path) as f:
        header = f.readline()
        if not header.startswith("%%MatrixMarket matrix array"):
            raise ValueError(f"{path}: not a dense Matrix Market file")
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        rows, cols = (int(t) for t in line.split())
        vals = np.loadtxt(f, ndmin=1)
    return vals.reshape(cols, rows).T
