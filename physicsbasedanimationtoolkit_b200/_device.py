"""Shared ctypes front-end of the two integrator classes: owns one ``vbdx_integrator`` handle."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _interleaved(a, dtype, nV, name):
    """3 x nV (reference convention) -> contiguous nV x 3, which is the column-major 3 x nV the
    C-ABI expects."""
    a = np.asarray(a, dtype=dtype)
    if a.shape != (3, nV):
        raise ValueError(f"{name} must be 3 x {nV}, got {a.shape}")  # gpu/impl/common/Eigen.cuh:28-35
    return np.ascontiguousarray(a.T)


class DeviceIntegrator:
    _dtype = np.float32

    @staticmethod
    def _describe(data, keep, *, device=-1, tile_iters=0, flags=0, colors=None, kernel_variant=0, ring_slots=0, consumer_warps=0,
                  ghosts=None, rest=False, n_colors=0):
        """Fill a ``vbdx_data_desc`` from a constructed ``Data`` (arrays are parked in ``keep``)."""
        L = _lib.lib()
        if data.x.size == 0:
            raise ValueError("Data.construct() must be called before creating an integrator")
        nV, nT = data.X.shape[1], data.E.shape[1]
        d = _lib.DataDesc()
        L.vbdx_data_desc_init(C.byref(d))

        def ptr(a, dtype, transpose=False):
            if a is None or np.size(a) == 0:
                return None
            a = np.asarray(a, dtype=dtype)
            a = np.ascontiguousarray(a.T if transpose else a)
            keep.append(a)
            return a.ctypes.data

        d.nV, d.nT = nV, nT
        # the reference uploads data.x (gpu/impl/vbd/Integrator.cu:27); the element rest data must come from X
        # (sim/vbd/Data.cpp:220-221), which differs from x only if the caller edited data.x after construct()
        d.X = ptr(data.X if rest else data.x, np.float64, True)
        d.E = ptr(data.E, np.int64, True)
        d.v = ptr(data.v, np.float64, True)
        d.aext = ptr(data.aext, np.float64, True)
        d.m = ptr(data.m, np.float64)
        d.rhoe = ptr(data.rhoe, np.float64)
        d.lame = ptr(data.lame, np.float64, True)
        d.dbc = ptr(data.dbc, np.int64)
        d.nDbc = int(np.size(data.dbc))
        cols = data.colors if colors is None else colors
        d.colors = ptr(cols, np.int64)
        d.ordering = int(data.vertex_coloring_ordering)
        d.selection = int(data.vertex_coloring_selection)
        d.strategy = int(data.strategy)
        d.acceleration = int(data.accelerator)
        d.omega_mode = int(getattr(data, "omega_mode", 0))
        d.material = int(getattr(data, "energy", 0))
        d.kD, d.detHZero, d.rho = float(data.kD), float(data.detH_zero), float(data.rho)
        d.B = ptr(data.B, np.int64)
        d.V = ptr(data.V, np.int64)
        d.nCV = int(np.size(data.V))
        d.F = ptr(data.F, np.int64, True)
        d.nF = int(data.F.shape[1]) if np.ndim(data.F) == 2 else 0
        d.muC, d.muF, d.epsv = float(data.muC), float(data.muF), float(data.epsv)
        d.active_set_update_frequency = int(data.active_set_update_frequency)
        d.device, d.tile_iters, d.flags = int(device), int(tile_iters), int(flags)
        d.kernel_variant, d.ring_slots = int(kernel_variant), int(ring_slots)
        d.consumer_warps = int(consumer_warps)
        d.window_size = int(getattr(data, "manderson", 5))
        d.n_colors = int(n_colors)
        d.nesterov_L, d.nesterov_start = float(getattr(data, "nesterov_L", 1.0)), int(getattr(data, "nesterov_start", 3))
        d.tr_eta, d.tr_tau, d.tr_curved = float(getattr(data, "eta", 0.2)), float(getattr(data, "tau", 2.0)), int(bool(getattr(data, "curved", True)))
        d.ghosts = ptr(ghosts, np.int64)
        d.nGhosts = 0 if ghosts is None else int(np.size(ghosts))
        return d

    def __init__(self, data, **tuning):
        L = _lib.lib()
        keep = []
        self._rest_differs = data.x.size > 0 and not np.array_equal(data.x, data.X)
        d = self._describe(data, keep, rest=self._rest_differs, **tuning)
        self._h = C.c_void_p()
        self._L = L
        _lib.check(L.vbdx_create(C.byref(d), C.byref(self._h)))
        self.nV, self.nT = int(d.nV), int(d.nT)
        self._ncv = int(d.nCV)
        self._strategy, self._kD, self._detH = int(data.strategy), float(data.kD), float(data.detH_zero)
        if self._rest_differs:
            self.x = data.x

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.vbdx_destroy(h)
            self._h = None

    # ---- stepping -------------------------------------------------------------------------
    def _step(self, dt, iterations, substeps):
        _lib.check(self._L.vbdx_step(self._h, float(dt), int(iterations), int(substeps)))

    PRE_STEP, POST_STEP = 1, 2

    def step_partial(self, sdt, k_begin, k_end, total_iterations, flags=0):
        """One slice of a substep (``vbdx_step_partial``): optional pre-step, iterations ``k_begin..k_end-1`` of a
        solve of ``total_iterations``, optional velocity update."""
        _lib.check(self._L.vbdx_step_partial(self._h, float(sdt), int(k_begin), int(k_end), int(total_iterations), int(flags)))

    def objective(self, xk, xtilde, dt, gradient=False):
        """``Integrator::ObjectiveFunction`` / ``ObjectiveFunctionGradient`` (sim/vbd/Integrator.cpp:138-200) at the
        given 3 x nV arrays, evaluated on the device in double precision.  Returns f, or (f, grad 3 x nV)."""
        xk = np.ascontiguousarray(np.asarray(xk, np.float64).T)
        xt = np.ascontiguousarray(np.asarray(xtilde, np.float64).T)
        if xk.shape != (self.nV, 3) or xt.shape != (self.nV, 3):
            raise ValueError(f"xk and xtilde must be 3 x {self.nV}")
        f = C.c_double(0.0)
        g = np.empty((self.nV, 3)) if gradient else None
        _lib.check(self._L.vbdx_objective(self._h, xk.ctypes.data, xt.ctypes.data, float(dt), C.byref(f),
                                          None if g is None else g.ctypes.data))
        return (f.value, np.ascontiguousarray(g.T)) if gradient else f.value

    def _traced_substeps(self, dt, iterations, substeps):
        """Generator over the iterates of one step: yields (substep, k, x, xtilde, sdt) before every sweep and once
        more (k = iterations) after the velocity update -- the points at which the reference records its trace."""
        sdt = dt / substeps
        f64 = np.float64
        for s in range(substeps):
            self.step_partial(sdt, 0, 0, iterations, self.PRE_STEP)
            xtilde = self._get("inertial_target").astype(f64)
            for k in range(iterations):
                yield s, k, self._get("positions").astype(f64), xtilde, sdt
                self.step_partial(sdt, k, k + 1, iterations, 0)
            self.step_partial(sdt, iterations, iterations, iterations, self.POST_STEP)
            yield s, iterations, self._get("positions").astype(f64), xtilde, sdt

    def step_async(self, dt, iterations, substeps=1):
        """Extension: enqueue a step on the handle's stream and return immediately."""
        _lib.check(self._L.vbdx_step_async(self._h, float(dt), int(iterations), int(substeps)))

    def synchronize(self):
        _lib.check(self._L.vbdx_synchronize(self._h))

    def use_stream(self, cuda_stream_ptr):
        """Extension: run on a caller-provided ``cudaStream_t`` (e.g. torch's current stream)."""
        _lib.check(self._L.vbdx_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    # ---- state ----------------------------------------------------------------------------
    def _sfx(self):
        return "f32" if self._dtype == np.float32 else "f64"

    _FIELDS = {"positions": 0, "velocities": 1, "external_acceleration": 2, "inertial_target": 3, "previous_positions": 4}

    def _get(self, what, out=None):
        """3 x nV array of ``what``.  ``out`` (C-contiguous 3 x nV of the integrator's dtype, ideally from
        ``pinned_empty``) receives the device-to-host copy directly."""
        if out is None:
            out = np.empty((3, self.nV), dtype=self._dtype)
        elif not (isinstance(out, np.ndarray) and out.shape == (3, self.nV) and out.dtype == self._dtype and out.flags.c_contiguous):
            raise ValueError(f"out must be a C-contiguous 3 x {self.nV} array of {np.dtype(self._dtype).name}")
        _lib.check(self._L.vbdx_get_vertex_field(self._h, self._FIELDS[what], int(self._dtype == np.float64), 1, out.ctypes.data, self.nV))
        return out

    def _set(self, what, a):
        a = np.asarray(a, dtype=self._dtype)
        if a.shape != (3, self.nV):
            raise ValueError(f"{what} must be 3 x {self.nV}, got {a.shape}")  # gpu/impl/common/Eigen.cuh:28-35
        # row-major (numpy default) and column-major (Eigen) inputs are both consumed in place
        rows = not a.flags.f_contiguous
        if rows and not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        _lib.check(self._L.vbdx_set_vertex_field(self._h, self._FIELDS[what], int(self._dtype == np.float64), int(rows), a.ctypes.data, self.nV))

    def positions(self, out=None):
        """``x`` with an optional preallocated (pinned) destination."""
        return self._get("positions", out)

    def _check_async_array(self, a):
        if not (isinstance(a, np.ndarray) and a.shape == (3, self.nV) and a.dtype == self._dtype and a.flags.c_contiguous):
            raise ValueError(f"need a C-contiguous 3 x {self.nV} array of {np.dtype(self._dtype).name} (ideally from pinned_empty)")

    def set_positions_async(self, a):
        """Enqueue the upload of ``a`` (page-locked, see ``host.pinned_empty``) on the integrator's stream without waiting;
        ``a`` must stay untouched until ``synchronize()``.  With ``step_async`` and ``positions_async`` a step costs the host one
        round trip instead of three."""
        self._check_async_array(a)
        _lib.check(self._L.vbdx_set_vertex_field_async(self._h, 0, int(self._dtype == np.float64), 1, a.ctypes.data, self.nV))

    def positions_async(self, out):
        """Enqueue the download of ``x`` into ``out`` (page-locked); valid after ``synchronize()``."""
        self._check_async_array(out)
        _lib.check(self._L.vbdx_get_vertex_field_async(self._h, 0, int(self._dtype == np.float64), 1, out.ctypes.data, self.nV))
        return out

    def velocities(self, out=None):
        return self._get("velocities", out)

    x = property(lambda s: s._get("positions"), lambda s, a: s._set("positions", a))
    v = property(lambda s: s._get("velocities"), lambda s, a: s._set("velocities", a))

    def _set_acceleration(self, a):
        self._set("external_acceleration", a)

    def _set_detH(self, zero):
        _lib.check(self._L.vbdx_set_detH_zero(self._h, float(zero)))
        self._detH = float(zero)

    def _set_kD(self, kD):
        _lib.check(self._L.vbdx_set_rayleigh_damping(self._h, float(kD)))
        self._kD = float(kD)

    def _set_strategy(self, strategy):
        _lib.check(self._L.vbdx_set_initialization_strategy(self._h, int(strategy)))
        self._strategy = int(strategy)

    def _set_line_search_guard(self, on):
        _lib.check(self._L.vbdx_set_line_search_guard(self._h, 1 if on else 0))

    line_search_guard = property(None, _set_line_search_guard,
                                 doc="Extension (write-only): guarded Newton step, see include/vbdx.h vbdx_set_line_search_guard")

    def _set_block_size(self, n):
        _lib.check(self._L.vbdx_set_block_size(self._h, int(n)))

    def _set_scene_bounding_box(self, lo, hi):
        lo = np.ascontiguousarray(lo, dtype=np.float32)
        hi = np.ascontiguousarray(hi, dtype=np.float32)
        _lib.check(self._L.vbdx_set_scene_bounding_box(self._h, lo.ctypes.data, hi.ctypes.data))

    # ---- introspection (extensions used by tests / bench) -----------------------------------
    @property
    def info(self):
        out = _lib.Info()
        _lib.check(self._L.vbdx_get_info(self._h, C.byref(out)))
        return {k: getattr(out, k) for k, _ in out._fields_}

    def adjacency(self):
        p = np.empty(self.nV + 1, np.int64)
        e = np.empty(4 * self.nT, np.int64)
        il = np.empty(4 * self.nT, np.int64)
        _lib.check(self._L.vbdx_get_adjacency(self._h, p.ctypes.data, e.ctypes.data, il.ctypes.data))
        return p, e, il

    def element_data(self):
        GP = np.empty(12 * self.nT)
        wg = np.empty(self.nT)
        m = np.empty(self.nV)
        _lib.check(self._L.vbdx_get_element_data(self._h, GP.ctypes.data, wg.ctypes.data, m.ctypes.data))
        return GP.reshape(3 * self.nT, 4).T.copy(), wg, m

    def colors(self):
        c = np.empty(self.nV, np.int64)
        _lib.check(self._L.vbdx_get_colors(self._h, c.ctypes.data))
        return c

    def trace_phases(self, iteration, dt, iterations, substeps=1):
        """Diagnostics (direct kernel): %globaltimer stamps [colour, CTA, 4] of one sweep iteration."""
        info = self.info
        n = info["nColors"] * info["gridBlocks"] * 12
        _lib.check(self._L.vbdx_debug_trace(self._h, int(iteration), None, 0))
        self._step(dt, iterations, substeps)
        out = np.zeros(n, dtype=np.uint64)
        _lib.check(self._L.vbdx_debug_trace(self._h, int(iteration), out.ctypes.data, n))
        return out.reshape(info["nColors"], info["gridBlocks"], 12)

    def contact_state(self):
        """Extension: (active[nCV], nn[nCV, 8], nActive) of the vertex-triangle active set."""
        ncv = self._ncv
        active = np.zeros(ncv, np.int32)
        nn = np.zeros((ncv, 8), np.int32)
        na = C.c_int64(0)
        _lib.check(self._L.vbdx_get_contact_state(self._h, active.ctypes.data, nn.ctypes.data, C.byref(na)))
        return active.astype(bool), nn, na.value

    # ---- domain decomposition plumbing (see dist.py) -----------------------------------------
    def internal_ids(self):
        out = np.empty(self.nV, np.int64)
        _lib.check(self._L.vbdx_get_internal_ids(self._h, out.ctypes.data))
        return out

    def ipc_handles(self):
        out = np.zeros(128, np.uint8)
        _lib.check(self._L.vbdx_dist_ipc_handles(self._h, out.ctypes.data))
        return out

    def dist_stats(self, reset=True):
        """Halo-exchange diagnostics (``vbdx_dist_stats``): late ghosts, ns polling, barriers that waited, ns."""
        out = np.zeros(4, np.uint32)
        _lib.check(self._L.vbdx_dist_stats(self._h, out.ctypes.data, int(reset)))
        return dict(late_ghosts=int(out[0]), ghost_poll_ns=int(out[1]), epoch_waits=int(out[2]), epoch_wait_ns=int(out[3]))

    def dist_connect(self, rank, world, all_handles, peer_nverts, peer_nghosts, send_local, send_peer, send_remote, recv_mask):
        all_handles = np.ascontiguousarray(all_handles, np.uint8)
        peer_nverts = np.ascontiguousarray(peer_nverts, np.int64)
        peer_nghosts = np.ascontiguousarray(peer_nghosts, np.int64)
        sl, sp, sr = (np.ascontiguousarray(a, np.int64) for a in (send_local, send_peer, send_remote))
        _lib.check(self._L.vbdx_dist_connect(self._h, int(rank), int(world), all_handles.ctypes.data, peer_nverts.ctypes.data, peer_nghosts.ctypes.data,
                                             sl.size, sl.ctypes.data, sp.ctypes.data, sr.ctypes.data, int(recv_mask)))
