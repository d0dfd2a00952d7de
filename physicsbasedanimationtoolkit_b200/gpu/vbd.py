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
