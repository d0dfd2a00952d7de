"""``pbat.sim.xpbd``: the XPBD problem description, mirroring bindings/pypbat/sim/xpbd/Data.cpp and
sim/xpbd/Data.h:19-104 / sim/xpbd/Data.cpp (same fluent ``with_*`` builder, field names, defaults and validation errors), and
the helpers of python/pbatoolkit/py/sim/xpbd.py.  The integrator runs on the GPU (``pbat.gpu.xpbd.Integrator``;
``pbat.sim.xpbd.Integrator`` is the same object behind a double-precision interface) -- there is no CPU integrator here."""
from __future__ import annotations

import enum

import numpy as np

from .. import graph as _graph
from .vbd import lame_coefficients


class Constraint(enum.IntEnum):
    """sim/xpbd/Enums.h:9-13"""
    StableNeoHookean = 0
    Collision = 1


def partition_mesh_constraints(X, E, ordering=_graph.GreedyColorOrderingStrategy.LargestDegree,
                               selection=_graph.GreedyColorSelectionStrategy.LeastUsed):
    """python/pbatoolkit/py/sim/xpbd.py:77-100: colours the constraint graph of a mesh whose constraints sit on its elements
    (two elements conflict when they share a vertex).  Returns ``(ptr, adj, GC)``: the partitions in compressed sparse
    format and the colour of every element."""
    GGT = _graph.mesh_dual_graph(E, np.asarray(X).shape[1])
    GC = _graph.greedy_color(GGT.indptr, GGT.indices, ordering=ordering, selection=selection)
    ptr, adj = _graph.map_to_adjacency(GC)
    return ptr, adj, GC


class Data:
    """Mirror of ``pbat::sim::xpbd::Data``.  Arrays are numpy, one column per particle / element."""

    def __init__(self):
        e = np.empty
        self.V = e(0, dtype=np.int64)
        self.F = e((3, 0), dtype=np.int64)
        self.T = e((4, 0), dtype=np.int64)
        self.BV = e(0, dtype=np.int64)
        self.x = e((3, 0))
        self.v = e((3, 0))
        self.aext = e((3, 0))
        self.minv = e(0)
        self.xt = e((3, 0))
        self.xb = e((3, 0))
        self.lame = e((2, 0))
        self.DmInv = e((3, 0))
        self.gammaSNH = e(0)
        self.muV = e(0)
        self.muS, self.muD = 0.3, 0.2
        self.active_set_update_frequency = 1
        self.alpha = [e(0), e(0)]
        self.beta = [e(0), e(0)]
        self.lambda_ = [e(0), e(0)]
        self.dbc = e(0, dtype=np.int64)
        self.Pptr, self.Padj = [], []
        self.SGptr, self.SGadj, self.Cptr, self.Cadj = [], [], [], []
        self._alpha_snh_given = False

    # ---- fluent builder (sim/xpbd/Data.cpp:20-100) ------------------------------------------
    def with_volume_mesh(self, V, E):
        self.x = np.asarray(V, dtype=np.float64).copy()
        self.xt = self.x.copy()
        self.T = np.asarray(E, dtype=np.int64).copy()
        if self.x.ndim != 2 or self.x.shape[0] != 3 or self.T.ndim != 2 or self.T.shape[0] != 4:
            raise ValueError("expected V 3 x |#verts| and E 4 x |#elements|")
        return self

    def with_surface_mesh(self, V, F):
        self.V = np.asarray(V, dtype=np.int64).reshape(-1).copy()
        self.F = np.asarray(F, dtype=np.int64).copy()
        return self

    def with_bodies(self, BV):
        self.BV = np.asarray(BV, dtype=np.int64).reshape(-1).copy()
        return self

    def with_velocity(self, v):
        self.v = np.asarray(v, dtype=np.float64).copy()
        return self

    def with_acceleration(self, aext):
        self.aext = np.asarray(aext, dtype=np.float64).copy()
        return self

    def with_mass_inverse(self, minv):
        self.minv = np.asarray(minv, dtype=np.float64).reshape(-1).copy()
        return self

    def with_elastic_material(self, lame):
        self.lame = np.asarray(lame, dtype=np.float64).copy()
        return self

    def with_collision_penalties(self, muV):
        self.muV = np.asarray(muV, dtype=np.float64).reshape(-1).copy()
        return self

    def with_friction_coefficients(self, muS, muD):
        self.muS, self.muD = float(muS), float(muD)
        return self

    def with_active_set_update_frequency(self, frequency):
        self.active_set_update_frequency = int(frequency)
        return self

    def with_damping(self, beta, constraint):
        self.beta[int(constraint)] = np.asarray(beta, dtype=np.float64).reshape(-1).copy()
        return self

    def with_compliance(self, alpha, constraint):
        self.alpha[int(constraint)] = np.asarray(alpha, dtype=np.float64).reshape(-1).copy()
        if int(constraint) == Constraint.StableNeoHookean:
            self._alpha_snh_given = True
        return self

    def with_partitions(self, Pptr, Padj):
        self.Pptr, self.Padj = list(np.asarray(Pptr).tolist()), list(np.asarray(Padj).tolist())
        return self

    def with_cluster_partitions(self, SGptr, SGadj, Cptr, Cadj):
        self.SGptr, self.SGadj = list(np.asarray(SGptr).tolist()), list(np.asarray(SGadj).tolist())
        self.Cptr, self.Cadj = list(np.asarray(Cptr).tolist()), list(np.asarray(Cadj).tolist())
        return self

    def with_dirichlet_constrained_vertices(self, dbc):
        self.dbc = np.asarray(dbc, dtype=np.int64).reshape(-1).copy()
        return self

    # ---- Data::Construct (sim/xpbd/Data.cpp:102-243) ----------------------------------------
    def construct(self, validate=True):
        x, T = self.x, self.T
        nV, nT = x.shape[1], T.shape[1]
        if self.v.size == 0:
            self.v = np.zeros_like(x)
        if self.aext.size == 0:
            self.aext = np.zeros_like(x)
            self.aext[-1] = -9.81
        if self.minv.size == 0:
            self.minv = np.full(nV, 1e-3)
        if self.BV.size == 0:
            self.BV = np.zeros(nV, dtype=np.int64)
        self.xb = x.copy()
        if self.dbc.size:
            self.minv[self.dbc] = 0.0
            self.v[:, self.dbc] = 0.0
            self.aext[:, self.dbc] = 0.0
        if self.lame.size == 0:
            mu, lam = lame_coefficients(1e6, 0.45)
            self.lame = np.empty((2, nT))
            self.lame[0], self.lame[1] = mu, lam
        snh, col = int(Constraint.StableNeoHookean), int(Constraint.Collision)
        Ds = np.stack([x[:, T[a]] - x[:, T[0]] for a in (1, 2, 3)], axis=2).transpose(1, 0, 2)     # nT x 3 x 3
        self.DmInv = np.ascontiguousarray(np.linalg.inv(Ds).transpose(1, 0, 2).reshape(3, 3 * nT))  # block t = columns 3t..3t+2
        vol = np.linalg.det(Ds) / 6.0
        if not self._alpha_snh_given:
            a = np.empty(2 * nT)
            a[0::2], a[1::2] = 1.0 / (self.lame[0] * vol), 1.0 / (self.lame[1] * vol)
            self.alpha[snh] = a
        self.gammaSNH = 1.0 + self.lame[0] / self.lame[1]
        if self.beta[snh].size == 0:
            self.beta[snh] = np.zeros(2 * nT)
        self.lambda_[snh] = np.zeros(2 * nT)
        if self.alpha[col].size == 0:
            self.alpha[col] = np.zeros(self.V.size)
        if self.beta[col].size == 0:
            self.beta[col] = np.zeros(self.V.size)
        self.lambda_[col] = np.zeros(self.V.size)
        if self.muV.size == 0:
            self.muV = np.ones(self.V.size)
        if validate:
            ok = (self.v.shape == x.shape and self.aext.shape == x.shape and self.xt.shape == x.shape and self.minv.size == nV and x.shape[0] == 3)
            if not ok:
                raise ValueError(f"x, v, aext and m must have same #columns={nV} as x, and 3 rows (except m)")
            if not (T.shape[0] == 4 and self.lame.shape == (2, nT)):
                raise ValueError(f"With #elements={nT}, expected T=4x{nT}, lame=2x{nT}")
            if not (self.BV.size == nV and self.muV.size == self.V.size):
                raise ValueError(f"Expected BV.size()={nV}, muV.size()={self.V.size}")
        return self


class Integrator:
    """``pbat.sim.xpbd.Integrator(data)``: ``step(dt, iterations, substeps)``, ``x`` / ``v`` (3 x |#particles|, float64), public
    ``data`` kept in sync like the reference whose Step mutates it (bindings/pypbat/sim/xpbd/Integrator.cpp)."""

    def __init__(self, data: Data, **tuning):
        from ..gpu.xpbd import Integrator as _Gpu

        self._impl = _Gpu(data, **tuning)
        self.data = data

    def step(self, dt, iterations, substeps=1):
        self._impl.step(dt, iterations, substeps)
        self.data.x = self.x
        self.data.v = self.v

    x = property(lambda s: s._impl._get("positions", np.float64), lambda s, a: s._impl._set("positions", a))
    v = property(lambda s: s._impl._get("velocities", np.float64), lambda s, a: s._impl._set("velocities", a))
