"""``pbat.sim.vbd``: the problem description (``Data``) and the double-precision-interface
integrator, mirroring bindings/pypbat/sim/vbd/{Data,Integrator}.cpp.

``Data`` is a host-side (numpy) mirror of ``pbat::sim::vbd::Data`` (sim/vbd/Data.h:27-246):
same fluent ``with_*`` builder, same field names, same defaults, same validation errors
(``ValueError`` where the reference throws ``std::invalid_argument``).  ``Integrator`` runs on
the GPU -- there is no CPU integrator in this package.
"""
from __future__ import annotations

import enum
import os

import numpy as np

from .. import graph as _graph
from .._device import DeviceIntegrator


class InitializationStrategy(enum.IntEnum):
    """sim/vbd/Enums.h:9-15"""
    Position = 0
    Inertia = 1
    KineticEnergyMinimum = 2
    AdaptiveVbd = 3
    AdaptivePbat = 4


class AccelerationStrategy(enum.IntEnum):
    """sim/vbd/Enums.h:21-28 (Python names: bindings/pypbat/sim/vbd/Data.cpp:24-30)"""
    Base = 0
    Chebyshev = 1
    Anderson = 2
    Nesterov = 3
    Broyden = 4
    TrustRegion = 5


class HyperElasticEnergy(enum.IntEnum):
    """Elastic energy density of the sweep (include/vbdx.h vbdx_material).  The reference's integrator
    instantiates ``physics::StableNeoHookeanEnergy<3>`` (sim/vbd/Integrator.cpp:120); ``SaintVenantKirchhoff``
    (physics/SaintVenantKirchhoffEnergy.h) is offered in its place as the north star asks."""
    StableNeoHookean = 0
    SaintVenantKirchhoff = 1


def lame_coefficients(Y, nu):
    """physics/HyperElasticity.cpp:6-11"""
    mu = Y / (2.0 * (1.0 + nu))
    lam = (Y * nu) / ((1.0 + nu) * (1.0 - 2.0 * nu))
    return mu, lam


def _cols(a, rows, dtype, name):
    a = np.asarray(a, dtype=dtype)
    if a.ndim != 2 or a.shape[0] != rows:
        raise ValueError(f"{name} must have {rows} rows, got shape {a.shape}")
    return a


class Data:
    """Mirror of ``pbat::sim::vbd::Data``.  Arrays are numpy, one column per vertex/element."""

    def __init__(self):
        e = np.empty
        self.X = e((3, 0))
        self.E = e((4, 0), dtype=np.int64)
        self.B = e(0, dtype=np.int64)
        self.V = e(0, dtype=np.int64)
        self.F = e((3, 0), dtype=np.int64)
        self.XVA = e(0)
        self.FA = e(0)
        self.x = e((3, 0))
        self.v = e((3, 0))
        self.aext = e((3, 0))
        self.m = e(0)
        self.xt = e((3, 0))
        self.xtilde = e((3, 0))
        self.vt = e((3, 0))
        self.wg = e(0)
        self._GP = e((4, 0))
        self.rhoe = e(0)
        self.lame = e((2, 0))
        self._GVG = (e(0, dtype=np.int64), e(0, dtype=np.int64), e(0, dtype=np.int64))
        self._lazy = set()   # fields construct() has not materialised yet (see GP, GVGp, GVGe, GVGilocal)
        self.muD = 1.0
        self.dbc = e(0, dtype=np.int64)
        self.vertex_coloring_ordering = _graph.GreedyColorOrderingStrategy.LargestDegree
        self.vertex_coloring_selection = _graph.GreedyColorSelectionStrategy.LeastUsed
        self.colors = e(0, dtype=np.int64)
        self.Pptr = e(0, dtype=np.int64)
        self.Padj = e(0, dtype=np.int64)
        # sim/vbd/Data.h:222-246
        self.strategy = InitializationStrategy.AdaptivePbat
        self.kD = 0.0
        self.muC = 1e6
        self.muF = 0.3
        self.epsv = 1e-3
        self.active_set_update_frequency = 1
        self.detH_zero = 1e-7
        self.accelerator = AccelerationStrategy.Base
        self.rho = 1.0
        self.manderson = 5
        self.nesterov_L = 1.0
        self.nesterov_start = 3
        self.eta = 0.2
        self.tau = 2.0
        self.curved = True
        # extension (not in the reference): which omega recurrence the Chebyshev solve uses,
        # 0 = as the reference evaluates it, 1 = textbook (include/vbdx.h vbdx_omega_mode)
        self.omega_mode = 0
        # extension: elastic energy (HyperElasticEnergy); the reference's only choice is the default
        self.energy = HyperElasticEnergy.StableNeoHookean

    # ---- fields materialised on demand ---------------------------------------------------
    @property
    def GP(self):
        """Shape function gradients at the quadrature points, 4 x 3|#elements| (fem/ShapeFunctions.h:267-297): rows 1..3 of
        element e's 4 x 3 block are the inverse of J = [X1-X0, X2-X0, X3-X0], row 0 is minus their sum."""
        if "GP" in self._lazy:
            X, E = self.X, self.E
            nT = E.shape[1]
            c0, c1, c2 = (X[:, E[a]] - X[:, E[0]] for a in (1, 2, 3))          # columns of J, 3 x nT each
            r0, r1, r2 = np.cross(c1, c2, axis=0), np.cross(c2, c0, axis=0), np.cross(c0, c1, axis=0)
            det = np.einsum("ij,ij->j", c0, r0)
            Jinv = np.stack([r0, r1, r2], axis=0) / det                           # [row, component, e]: rows of J^-1
            G = np.concatenate([-Jinv.sum(axis=0, keepdims=True), Jinv], axis=0)  # 4 x 3 x nT
            self._GP = np.ascontiguousarray(G.transpose(0, 2, 1).reshape(4, 3 * nT))
            self._lazy.discard("GP")
        return self._GP

    @GP.setter
    def GP(self, a):
        self._GP = a
        self._lazy.discard("GP")

    def _adjacency(self):
        if "GVG" in self._lazy:
            # vertex -> tet adjacency, ascending element id per vertex (sim/vbd/Data.cpp:223-226)
            E, nV = self.E, self.X.shape[1]
            flat = E.T.reshape(-1)                       # entry k = 4 e + ilocal
            order = np.argsort(flat, kind="stable")
            ptr = np.concatenate([[0], np.cumsum(np.bincount(flat, minlength=nV))]).astype(np.int64)
            self._GVG = (ptr, (order // 4).astype(np.int64), (order % 4).astype(np.int64))
            self._lazy.discard("GVG")
        return self._GVG

    GVGp = property(lambda s: s._adjacency()[0])
    GVGe = property(lambda s: s._adjacency()[1])
    GVGilocal = property(lambda s: s._adjacency()[2])

    # ---- fluent builder (sim/vbd/Data.cpp:20-177) ----------------------------------------
    def with_volume_mesh(self, X, T):
        self.X = _cols(X, 3, np.float64, "X").copy()
        self.E = _cols(T, 4, np.int64, "T").copy()
        if self.B.size == 0:
            self.B = np.ones(self.X.shape[1], dtype=np.int64)
        return self

    def with_surface_mesh(self, V, F):
        # vertex areas XVA and triangle areas FA as sim/vbd/Data.cpp:34-54
        self.V = np.asarray(V, dtype=np.int64).reshape(-1).copy()
        self.F = _cols(F, 3, np.int64, "F").copy()
        X = self.X
        AB = X[:, self.F[1]] - X[:, self.F[0]]
        AC = X[:, self.F[2]] - X[:, self.F[0]]
        dbl = np.linalg.norm(np.cross(AB, AC, axis=0), axis=0)
        self.XVA = np.zeros(X.shape[1])
        for r in range(3):
            np.add.at(self.XVA, self.F[r], dbl / 6.0)
        self.FA = dbl / 2.0
        return self

    def with_bodies(self, B):
        self.B = np.asarray(B, dtype=np.int64).reshape(-1).copy()
        return self

    def with_velocity(self, v):
        self.v = _cols(v, 3, np.float64, "v").copy()
        return self

    def with_acceleration(self, a):
        self.aext = _cols(a, 3, np.float64, "a").copy()
        return self

    def with_material(self, rhoe, mue, lambdae):
        self.rhoe = np.asarray(rhoe, dtype=np.float64).reshape(-1).copy()
        self.lame = np.stack([np.asarray(mue, np.float64).reshape(-1),
                              np.asarray(lambdae, np.float64).reshape(-1)])
        return self

    def with_hyper_elastic_energy(self, energy):
        """Extension: choose the elastic energy density (``HyperElasticEnergy``)."""
        self.energy = HyperElasticEnergy(energy)
        return self

    def with_dirichlet_vertices(self, dbc, muD=1.0, input_sorted=True):
        self.dbc = np.asarray(dbc, dtype=np.int64).reshape(-1).copy()
        self.muD = float(muD)
        if not input_sorted:
            self.dbc.sort()
        return self

    def with_vertex_coloring_strategy(self, ordering, selection):
        self.vertex_coloring_ordering = _graph.GreedyColorOrderingStrategy(ordering)
        self.vertex_coloring_selection = _graph.GreedyColorSelectionStrategy(selection)
        return self

    def with_initialization_strategy(self, strategy):
        self.strategy = InitializationStrategy(strategy)
        return self

    def with_rayleigh_damping(self, kD):
        self.kD = float(kD)
        return self

    def with_contact_parameters(self, muC, muF, epsv):
        self.muC, self.muF, self.epsv = float(muC), float(muF), float(epsv)
        return self

    def with_active_set_update_frequency(self, frequency):
        self.active_set_update_frequency = int(frequency)
        return self

    def with_hessian_determinant_zero(self, zero):
        self.detH_zero = float(zero)
        return self

    def with_chebyshev_acceleration(self, rho):
        self.rho = float(rho)
        self.accelerator = AccelerationStrategy.Chebyshev
        return self

    def with_anderson_acceleration(self, window_size):
        self.manderson = int(window_size)
        self.accelerator = AccelerationStrategy.Anderson
        return self

    def with_broyden_acceleration(self, window_size):
        self.manderson = int(window_size)
        self.accelerator = AccelerationStrategy.Broyden
        return self

    def with_nesterov_acceleration(self, L, start):
        self.nesterov_L, self.nesterov_start = float(L), int(start)
        self.accelerator = AccelerationStrategy.Nesterov
        return self

    def with_trust_region_acceleration(self, eta, tau, curved):
        self.eta, self.tau, self.curved = float(eta), float(tau), bool(curved)
        self.accelerator = AccelerationStrategy.TrustRegion
        return self

    # ---- Data::Construct (sim/vbd/Data.cpp:179-308) --------------------------------------
    def construct(self, validate=True, coloring_device=None):
        """``Data::Construct`` (sim/vbd/Data.cpp:179-243).  ``coloring_device=<cuda ordinal>`` (extension) computes the vertex
        colouring on that GPU instead of the host -- same colours; for the FirstAvailable selection only."""
        X, E = self.X, self.E
        nV, nT = X.shape[1], E.shape[1]
        if nT and (E.min() < 0 or E.max() >= nV):
            raise ValueError("element index out of range")
        self.x = X.copy()
        if self.xt.size == 0:
            self.xt = self.x.copy()
        if self.v.size == 0:
            self.v = np.zeros_like(self.x)
        if self.aext.size == 0:
            self.aext = np.zeros_like(self.x)
            self.aext[2] = -9.81
        self.xtilde = np.zeros_like(self.x)
        self.vt = np.zeros_like(self.x)
        if self.lame.size == 0:
            mu, lam = lame_coefficients(1e6, 0.45)
            self.lame = np.empty((2, nT))
            self.lame[0], self.lame[1] = mu, lam
        if self.rhoe.size == 0:
            self.rhoe = np.full(nT, 1e3)
        # quadrature weights (fem/MeshQuadrature.h:75-88) and lumped mass (fem/Mass.h:791-830) for linear tets.  The shape
        # function gradients GP and the vertex -> tet adjacency GVG* are large and the device recomputes both (the north
        # star: CSR built on device), so they are materialised on first access (properties below).
        e1, e2, e3 = (X[:, E[a]] - X[:, E[0]] for a in (1, 2, 3))
        det = np.einsum("ij,ij->j", e1, np.cross(e2, e3, axis=0))
        if np.any(det <= 1e-10):
            raise ValueError("inverted or degenerate tetrahedron in the rest mesh")  # fem/Jacobian.h:68-80
        self.wg = det / 6.0
        self.m = np.bincount(E.reshape(-1), weights=np.tile(self.rhoe * self.wg / 4.0, 4), minlength=nV)
        self._lazy = {"GP", "GVG"}
        # colouring and partitions (sim/vbd/Data.cpp:228-231)
        self.colors = _graph.mesh_greedy_color(E, nV, self.vertex_coloring_ordering,
                                               self.vertex_coloring_selection, device=coloring_device)
        nC = int(self.colors.max()) + 1 if nV else 0
        keep = np.ones(nV, dtype=bool)
        if self.dbc.size:  # sim/vbd/Data.cpp:236-243
            if self.dbc.min() < 0 or self.dbc.max() >= nV:
                raise ValueError("Dirichlet vertex index out of range")
            self.v[:, self.dbc] = 0.0
            self.aext[:, self.dbc] = 0.0
            keep[self.dbc] = False
        verts = np.flatnonzero(keep)
        part = np.argsort(self.colors[verts], kind="stable")
        self.Padj = verts[part].astype(np.int64)
        self.Pptr = np.concatenate([[0], np.cumsum(np.bincount(self.colors[verts], minlength=nC))]).astype(np.int64)
        if validate:  # sim/vbd/Data.cpp:245-306
            ok = (self.xt.shape == self.x.shape and self.v.shape == self.x.shape and
                  self.aext.shape == self.x.shape and self.m.size == nV and self.B.size == nV and
                  self.x.shape[0] == 3)
            if not ok:
                raise ValueError(
                    f"x, v, aext, m and B must have same #columns={nV} as x, and 3 rows (except m and B)")
            A = AccelerationStrategy
            if self.accelerator == A.Chebyshev and not (0 < self.rho < 1):
                raise ValueError("Expected 0 < rho < 1")
            if self.accelerator in (A.Anderson, A.Broyden) and self.manderson < 1:
                raise ValueError("Expected m > 0")
            if self.accelerator == A.Nesterov:
                if self.nesterov_L <= 0:
                    raise ValueError("Expected L > 0")
                if self.nesterov_start < 0:
                    raise ValueError("Expected start >= 0")
            if self.accelerator == A.TrustRegion:
                if self.eta < 0:
                    raise ValueError("Expected eta >= 0")
                if self.tau <= 1:
                    raise ValueError("Expected tau > 1")
        return self


class Integrator(DeviceIntegrator):
    """``pbat.sim.vbd.Integrator`` (bindings/pypbat/sim/vbd/Integrator.cpp:30-90): positions and
    velocities cross the boundary as float64 3 x nV arrays.  Dispatch on ``data.accelerator``
    happens at construction like the reference's factory (all of ``AccelerationStrategy``: Base, Chebyshev, Anderson,
    Nesterov, Broyden, TrustRegion)."""

    _dtype = np.float64

    def __init__(self, data: Data, **tuning):
        super().__init__(data, **tuning)
        self.data = data
        self._trace = None

    def step(self, dt, iterations, substeps=1):
        if self._trace is None:
            self._step(dt, iterations, substeps)
        else:
            self._export_trace(dt, iterations, substeps)
        # keep the public `data` member in sync, like the reference whose Step mutates data.x / data.v
        self.data.x = self.x
        self.data.v = self.v

    def trace_next_step(self, path=".", t=-1):
        """``Integrator::TraceNextStep`` (sim/vbd/Integrator.cpp:47-52): the next ``step`` records, before every
        sweep and after the velocity update of each substep, the objective, its gradient and the iterate, and
        writes ``<path>/<t>.<substep>.{f,grad,x}.mtx`` (``ExportTrace``, sim/vbd/Integrator.cpp:202-224)."""
        self._trace = (str(path), int(t))

    def _export_trace(self, dt, iterations, substeps):
        from .. import mtx

        path, t = self._trace
        self._trace = None  # the reference clears mTraceIterates at the end of Step
        f, G, X = [], [], []
        for s, k, x, xtilde, sdt in self._traced_substeps(dt, iterations, substeps):
            fk, gk = self.objective(x, xtilde, sdt, gradient=True)
            f.append(fk), G.append(gk.T.reshape(-1)), X.append(x.T.reshape(-1))
            if k == iterations:
                mtx.save_dense(os.path.join(path, f"{t}.{s}.f.mtx"), np.asarray(f).reshape(-1, 1))
                mtx.save_dense(os.path.join(path, f"{t}.{s}.grad.mtx"), np.stack(G, axis=1))
                mtx.save_dense(os.path.join(path, f"{t}.{s}.x.mtx"), np.stack(X, axis=1))
                f, G, X = [], [], []

    def objective_function(self, xk, xtilde, dt):
        """``Integrator::ObjectiveFunction`` (sim/vbd/Integrator.h:45-51)."""
        return self.objective(xk, xtilde, dt)

    def objective_function_gradient(self, xk, xtilde, dt):
        """``Integrator::ObjectiveFunctionGradient`` (sim/vbd/Integrator.h:52-58): 3|#verts| vector."""
        return self.objective(xk, xtilde, dt, gradient=True)[1].T.reshape(-1)

    strategy = property(lambda s: InitializationStrategy(s._strategy), lambda s, v: s._set_strategy(v))
    kD = property(lambda s: s._kD, lambda s, v: s._set_kD(v))
    detH_residual = property(lambda s: s._detH, lambda s, v: s._set_detH(v))
