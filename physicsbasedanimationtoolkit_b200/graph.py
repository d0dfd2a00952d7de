"""``pbat.graph`` names used by the VBD path (bindings/pypbat/graph/Color.cpp:15-27)."""
from __future__ import annotations

import enum

import numpy as np

from . import _lib


class GreedyColorOrderingStrategy(enum.IntEnum):
    Natural = 0
    SmallestDegree = 1
    LargestDegree = 2


class GreedyColorSelectionStrategy(enum.IntEnum):
    LeastUsed = 0
    FirstAvailable = 1


def mesh_greedy_color(T, n_vertices, ordering=GreedyColorOrderingStrategy.LargestDegree,
                      selection=GreedyColorSelectionStrategy.LeastUsed):
    """Greedy colouring of the primal graph of a tet mesh ``T`` (4 x nT), exactly as
    ``Data::Construct`` obtains it (sim/vbd/Data.cpp:228-231 -> graph/Mesh.h:116-123,
    graph/Color.h:45-135).  Host-only; needs no GPU."""
    T = np.ascontiguousarray(np.asarray(T, dtype=np.int64).T)  # column-major 4 x nT
    out = np.empty(n_vertices, dtype=np.int64)
    _lib.check(_lib.lib().vbdx_greedy_color(n_vertices, T.shape[0], T.ctypes.data, int(ordering),
                                            int(selection), out.ctypes.data))
    return out
