"""``pbat.gpu.vbd.Integrator`` (bindings/pypbat/gpu/vbd/Integrator.cpp:26-104): float32 interface,
write-only tuning properties, ``step(dt=0.01, iterations=20, substeps=1)``."""
from __future__ import annotations

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

    def traced_step(self, dt=0.01, iterations=20, substeps=1, t=0, dir="."):
        raise NotImplementedError("iterate tracing is not implemented (SURVEY.md section 8f, rank 3)")

    a = _write_only(lambda s, a: s._set_acceleration(a))
    detH_residual = _write_only(lambda s, v: s._set_detH(v))
    kD = _write_only(lambda s, v: s._set_kD(v))
    strategy = _write_only(lambda s, v: s._set_strategy(v))
    gpu_block_size = _write_only(lambda s, v: s._set_block_size(v))
    scene_bounding_box = _write_only(lambda s, box: s._set_scene_bounding_box(box[0], box[1]))
