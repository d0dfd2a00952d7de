"""``pbat.gpu.vbd.Integrator`` (bindings/pypbat/gpu/vbd/Integrator.cpp:26-104): float32 interface,
write-only tuning properties, ``step(dt=0.01, iterations=20, substeps=1)``."""
from __future__ import annotations

import os

import numpy as np

from .._device import DeviceIntegrator


def _write_only(setter):
    def getter(self):
        raise AttributeError("write-only property")
    return property(getter, setter)


class Integrator(DeviceIntegrator):
    _dtype = np.float32

    def step(self, dt=0.01, iterations=20, substeps=1):
        self._step(dt, iterations, substeps)

    def __init__(self, data, **tuning):
        super().__init__(data, **tuning)
        self._trace_static = (np.asarray(data.E), np.asarray(data.lame))

    def traced_step(self, dt=0.01, iterations=20, substeps=1, t=0, dir="."):
        """``Integrator::TracedStep`` (gpu/impl/vbd/Integrator.cu:105-148,284-301): one step whose iterates are
        written as dense Matrix Market files into ``dir`` -- on the first frame the static problem data (``T``,
        ``M``, ``wg``, ``GP``, ``lame``), per substep ``xtilde.t.<t>.s.<s>.mtx`` and, before every sweep and after
        the last one, ``x.t.<t>.s.<s>.k.<k>.mtx`` (all |#verts| x 3)."""
        from .. import mtx

        if t == 0:
            E, lame = self._trace_static
            GP, wg, m = self.element_data()
            mtx.save_dense(os.path.join(dir, "T.mtx"), E.T)
            mtx.save_dense(os.path.join(dir, "M.mtx"), m.astype(np.float32).reshape(-1, 1))
            mtx.save_dense(os.path.join(dir, "wg.mtx"), wg.astype(np.float32).reshape(-1, 1))
            mtx.save_dense(os.path.join(dir, "GP.mtx"), GP.astype(np.float32))
            mtx.save_dense(os.path.join(dir, "lame.mtx"), lame.T.astype(np.float32))
        for s, k, x, xtilde, _ in self._traced_substeps(dt, iterations, substeps):
            if k == 0:
                mtx.save_dense(os.path.join(dir, f"xtilde.t.{t}.s.{s}.mtx"), xtilde.T.astype(np.float32))
            mtx.save_dense(os.path.join(dir, f"x.t.{t}.s.{s}.k.{k}.mtx"), x.T.astype(np.float32))

    a = _write_only(lambda s, a: s._set_acceleration(a))
    detH_residual = _write_only(lambda s, v: s._set_detH(v))
    kD = _write_only(lambda s, v: s._set_kD(v))
    strategy = _write_only(lambda s, v: s._set_strategy(v))
    gpu_block_size = _write_only(lambda s, v: s._set_block_size(v))
    scene_bounding_box = _write_only(lambda s, box: s._set_scene_bounding_box(box[0], box[1]))


class BatchIntegrator(Integrator):
    """Extension: independent scenes stepped together (``vbdx_create_batch``; SURVEY.md 8e, BASELINE configs[4]).

    ``datas`` is a sequence of constructed ``pbat.sim.vbd.Data`` that agree in the solver settings.  The object is an
    ``Integrator`` over the concatenation of the scenes: ``x`` / ``v`` are 3 x sum(nV); ``offsets[s]`` is the first
    vertex of scene ``s`` and ``scene(a, s)`` slices an array.  A scene in a batch evolves bit-identically to the
    scene stepped alone."""

    def __init__(self, datas, **tuning):
        import ctypes as C

        from .. import _lib

        datas = list(datas)
        if not datas:
            raise ValueError("a batch needs at least one scene")
        L = _lib.lib()
        keep = []
        descs = (_lib.DataDesc * len(datas))()
        for i, data in enumerate(datas):
            if data.x.size and not np.array_equal(data.x, data.X):
                raise ValueError("batch scenes must start from their rest positions (set x after construction)")
            descs[i] = self._describe(data, keep, **tuning)
        self._h = C.c_void_p()
        self._L = L
        _lib.check(L.vbdx_create_batch(descs, len(datas), C.byref(self._h)))
        self.offsets = np.zeros(len(datas) + 1, np.int64)
        _lib.check(L.vbdx_batch_offsets(self._h, None, self.offsets.ctypes.data))
        self.nV, self.nT = int(self.offsets[-1]), int(sum(d.E.shape[1] for d in datas))
        self.n_scenes = len(datas)
        self._ncv = 0
        self._rest_differs = False
        d0 = datas[0]
        self._strategy, self._kD, self._detH = int(d0.strategy), float(d0.kD), float(d0.detH_zero)
        self._trace_static = None

    def scene(self, a, s):
        """Columns of a 3 x sum(nV) array that belong to scene ``s``."""
        return a[:, self.offsets[s]:self.offsets[s + 1]]


class MultiGpuBatchIntegrator:
    """Extension: independent scenes sharded over several GPUs of one node (SURVEY.md 8e second row, BASELINE configs[4]:
    4096 scenes over 8 GPUs).  The scenes are split into contiguous blocks, one ``BatchIntegrator`` (one persistent launch per
    step) per device; a step enqueues all devices' launches and then waits for all of them -- there is no communication.
    ``x`` / ``v`` are 3 x sum(nV) over all scenes in the caller's scene order; scene ``s`` lives on ``device_of[s]``."""

    def __init__(self, datas, devices=None, **tuning):
        from .. import _lib

        datas = list(datas)
        if devices is None:
            devices = list(range(max(1, _lib.lib().vbdx_device_count())))
        devices = list(devices)[:max(1, len(datas))]
        if not devices:
            raise ValueError("no device")
        bounds = [(len(datas) * i) // len(devices) for i in range(len(devices) + 1)]
        self.devices = devices
        self.parts = [BatchIntegrator(datas[bounds[i]:bounds[i + 1]], device=dev, **tuning) for i, dev in enumerate(devices)]
        self.device_of = np.concatenate([np.full(bounds[i + 1] - bounds[i], dev) for i, dev in enumerate(devices)])
        self.n_scenes = len(datas)
        offs = [0]
        for p in self.parts:
            offs.extend((offs[-1] - p.offsets[0] + p.offsets[1:]).tolist())
        self.offsets = np.asarray(offs, np.int64)
        self._vbounds = np.concatenate([[0], np.cumsum([p.nV for p in self.parts])])
        self.nV = int(self._vbounds[-1])

    def step(self, dt=0.01, iterations=20, substeps=1):
        for p in self.parts:
            p.step_async(dt, iterations, substeps)
        for p in self.parts:
            p.synchronize()

    def _gather(self, what):
        return np.concatenate([getattr(p, what) for p in self.parts], axis=1)

    def _scatter(self, what, a):
        a = np.asarray(a)
        if a.shape != (3, self.nV):
            raise ValueError(f"{what} must be 3 x {self.nV}, got {a.shape}")
        for i, p in enumerate(self.parts):
            setattr(p, what, np.ascontiguousarray(a[:, self._vbounds[i]:self._vbounds[i + 1]]))

    x = property(lambda s: s._gather("x"), lambda s, a: s._scatter("x", a))
    v = property(lambda s: s._gather("v"), lambda s, a: s._scatter("v", a))

    def scene(self, a, s):
        """Columns of a 3 x sum(nV) array that belong to scene ``s``."""
        return a[:, self.offsets[s]:self.offsets[s + 1]]

    @property
    def info(self):
        return [p.info for p in self.parts]
