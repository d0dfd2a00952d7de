"""``pbat.gpu.xpbd.Integrator`` (bindings/pypbat/gpu/xpbd/Integrator.cpp:26-80): ``step(dt=0.01, iterations=10, substeps=5)``,
``x`` (read/write, float32), ``v`` (the reference exposes the setter only; reading works here too), ``set_compliance(alpha,
constraint_type)``, write-only ``mu = (muS, muK)`` and ``scene_bounding_box = (min, max)``."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib


def _write_only(setter):
    def getter(self):
        raise AttributeError("write-only property")
    return property(getter, setter)


class Integrator:
    def __init__(self, data, device=-1):
        L = _lib.lib()
        if data.gammaSNH.size == 0:
            raise ValueError("Data.construct() must be called before creating an integrator")
        if len(data.Pptr) < 2:
            raise ValueError("XPBD needs constraint partitions (Data.with_partitions)")
        keep = []

        def ptr(a, dtype, transpose=False):
            if a is None or np.size(a) == 0:
                return None
            a = np.asarray(a, dtype=dtype)
            a = np.ascontiguousarray(a.T if transpose else a)
            keep.append(a)
            return a.ctypes.data

        d = _lib.XpbdDesc()
        L.vbdx_xpbd_desc_init(C.byref(d))
        self.nV, self.nT = data.x.shape[1], data.T.shape[1]
        d.nV, d.nT = self.nV, self.nT
        d.X, d.T = ptr(data.x, np.float64, True), ptr(data.T, np.int64, True)
        d.v, d.aext = ptr(data.v, np.float64, True), ptr(data.aext, np.float64, True)
        d.minv, d.lame = ptr(data.minv, np.float64), ptr(data.lame, np.float64, True)
        d.dbc, d.nDbc = ptr(data.dbc, np.int64), int(np.size(data.dbc))
        d.Pptr, d.Padj, d.nPartitions = ptr(data.Pptr, np.int64), ptr(data.Padj, np.int64), len(data.Pptr) - 1
        if len(data.SGptr) > 1:
            d.SGptr, d.SGadj = ptr(data.SGptr, np.int64), ptr(data.SGadj, np.int64)
            d.Cptr, d.Cadj = ptr(data.Cptr, np.int64), ptr(data.Cadj, np.int64)
            d.nClusterPartitions = len(data.SGptr) - 1
        d.alphaSNH, d.betaSNH = ptr(data.alpha[0], np.float64), ptr(data.beta[0], np.float64)
        d.BV = ptr(data.BV, np.int64)
        d.V, d.nCV = ptr(data.V, np.int64), int(np.size(data.V))
        d.F = ptr(data.F, np.int64, True)
        d.nF = int(data.F.shape[1]) if np.ndim(data.F) == 2 else 0
        if d.nF == 0:
            d.nCV = 0
        d.muV, d.alphaC, d.betaC = ptr(data.muV, np.float64), ptr(data.alpha[1], np.float64), ptr(data.beta[1], np.float64)
        d.muS, d.muD = float(data.muS), float(data.muD)
        d.active_set_update_frequency = int(data.active_set_update_frequency)
        d.device = int(device)
        self._ncv = int(d.nCV)
        self._h = C.c_void_p()
        self._L = L
        _lib.check(L.vbdx_xpbd_create(C.byref(d), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.vbdx_xpbd_destroy(h)
            self._h = None

    def step(self, dt=0.01, iterations=10, substeps=5):
        _lib.check(self._L.vbdx_xpbd_step(self._h, float(dt), int(iterations), int(substeps)))

    def _get(self, what, dtype=np.float32):
        out = np.empty((self.nV, 3), np.float64)
        fn = self._L.vbdx_xpbd_get_positions if what == "positions" else self._L.vbdx_xpbd_get_velocities
        _lib.check(fn(self._h, out.ctypes.data, self.nV))
        return np.ascontiguousarray(out.T).astype(dtype)

    def _set(self, what, a):
        a = np.asarray(a, dtype=np.float64)
        if a.shape != (3, self.nV):
            raise ValueError(f"{what} must be 3 x {self.nV}, got {a.shape}")
        a = np.ascontiguousarray(a.T)
        fn = {"positions": self._L.vbdx_xpbd_set_positions, "velocities": self._L.vbdx_xpbd_set_velocities,
              "external_acceleration": self._L.vbdx_xpbd_set_external_acceleration}[what]
        _lib.check(fn(self._h, a.ctypes.data, self.nV))

    x = property(lambda s: s._get("positions"), lambda s, a: s._set("positions", a))
    v = property(lambda s: s._get("velocities"), lambda s, a: s._set("velocities", a))
    a = _write_only(lambda s, a: s._set("external_acceleration", a))

    def set_compliance(self, alpha, constraint_type):
        alpha = np.ascontiguousarray(np.asarray(alpha, dtype=np.float64).reshape(-1))
        _lib.check(self._L.vbdx_xpbd_set_compliance(self._h, int(constraint_type), alpha.ctypes.data, alpha.size))

    mu = _write_only(lambda s, mu: _lib.check(s._L.vbdx_xpbd_set_friction_coefficients(s._h, float(mu[0]), float(mu[1]))))

    def _set_box(self, box):
        lo, hi = (np.ascontiguousarray(b, dtype=np.float32) for b in box)
        _lib.check(self._L.vbdx_xpbd_set_scene_bounding_box(self._h, lo.ctypes.data, hi.ctypes.data))

    scene_bounding_box = _write_only(_set_box)

    @property
    def info(self):
        out = np.zeros(8, np.int64)
        _lib.check(self._L.vbdx_xpbd_get_info(self._h, out.ctypes.data))
        return dict(nV=int(out[0]), nT=int(out[1]), nPartitions=int(out[2]), gridBlocks=int(out[3]), kernelLaunches=int(out[4]),
                    deviceBytes=int(out[5]), lastStepMs=out[6] / 1e6, nCV=int(out[7]))

    def contact_state(self):
        active = np.zeros(self._ncv, np.int32)
        nn = np.zeros((self._ncv, 8), np.int32)
        na = C.c_int64(0)
        _lib.check(self._L.vbdx_xpbd_get_contact_state(self._h, active.ctypes.data, nn.ctypes.data, C.byref(na)))
        return active.astype(bool), nn, na.value
