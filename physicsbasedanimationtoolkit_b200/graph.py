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
                      selection=GreedyColorSelectionStrategy.LeastUsed, device=None):
    """Greedy colouring of the primal graph of a tet mesh ``T`` (4 x nT), exactly as
    ``Data::Construct`` obtains it (sim/vbd/Data.cpp:228-231 -> graph/Mesh.h:116-123,
    graph/Color.h:45-135).  On the host by default (needs no GPU); ``device=<cuda ordinal>`` computes the
    same colouring on that GPU -- for ``FirstAvailable`` only (``LeastUsed`` is inherently sequential:
    ``NotImplementedError``)."""
    T = np.ascontiguousarray(np.asarray(T, dtype=np.int64).T)  # column-major 4 x nT
    out = np.empty(n_vertices, dtype=np.int64)
    if device is not None:
        _lib.check(_lib.lib().vbdx_greedy_color_device(n_vertices, T.shape[0], T.ctypes.data, int(ordering), int(selection), int(device),
                                                       out.ctypes.data, None))
        return out
    _lib.check(_lib.lib().vbdx_greedy_color(n_vertices, T.shape[0], T.ctypes.data, int(ordering),
                                            int(selection), out.ctypes.data))
    return out


def greedy_color(ptr, adj, ordering=GreedyColorOrderingStrategy.LargestDegree, selection=GreedyColorSelectionStrategy.LeastUsed):
    """``pbat.graph.greedy_color`` (bindings/pypbat/graph/Color.cpp:28-60): greedy colouring of a graph in compressed sparse
    format, graph/Color.h:45-135.  Host-only."""
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    adj = np.ascontiguousarray(adj, dtype=np.int64)
    n = ptr.size - 1
    out = np.empty(n, dtype=np.int64)
    _lib.check(_lib.lib().vbdx_graph_greedy_color(n, ptr.ctypes.data, adj.ctypes.data, int(ordering), int(selection), out.ctypes.data))
    return out


def mesh_adjacency_matrix(C, n=-1):
    """Element -> vertex incidence (graph/Mesh.h:42-76) as a scipy CSC matrix: G[v, c] = 1 for every vertex v of element c."""
    import scipy.sparse as sp

    C = np.asarray(C, dtype=np.int64)
    n = int(C.max()) + 1 if n < 0 else int(n)
    cols = np.repeat(np.arange(C.shape[1]), C.shape[0])
    return sp.csc_matrix((np.ones(C.size, np.int64), (C.T.reshape(-1), cols)), shape=(n, C.shape[1]))


def mesh_primal_graph(C, n=-1):
    """``pbat.graph.mesh_primal_graph`` (graph/Mesh.h:116-123): G G^T, vertices adjacent through an element (self loops kept)."""
    G = mesh_adjacency_matrix(C, n)
    return (G @ G.T).tocsc()


def mesh_dual_graph(C, n=-1, flags=0b111):
    """``pbat.graph.mesh_dual_graph`` (graph/Mesh.h:137-178): G^T G, elements adjacent through >= 1 shared vertex; ``flags``
    keeps pairs sharing exactly 1 vertex (0b001), an edge (0b010), a face (0b100)."""
    G = mesh_adjacency_matrix(C, n)
    GT = (G.T @ G).tocsc()
    if flags != 0b111:
        keep = np.zeros_like(GT.data, dtype=bool)
        for bit, shared in ((0b001, 1), (0b010, 2), (0b100, 3)):
            if flags & bit:
                keep |= GT.data == shared
        GT.data = np.where(keep, GT.data, 0)
        GT.eliminate_zeros()
    return GT


def map_to_adjacency(p, n=-1):
    """``pbat.graph.map_to_adjacency`` (graph/Adjacency.h:171-181): (ptr, adj) of the partitions of a map vertex -> partition;
    members keep their order (stable)."""
    p = np.asarray(p, dtype=np.int64)
    n = int(p.max()) + 1 if n < 0 else int(n)
    ptr = np.concatenate([[0], np.cumsum(np.bincount(p, minlength=n))]).astype(np.int64)
    return ptr, np.argsort(p, kind="stable").astype(np.int64)
