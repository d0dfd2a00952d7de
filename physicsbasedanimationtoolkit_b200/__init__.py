"""B200-native Vertex Block Descent integrator behind the PBAT Python surface.

``import physicsbasedanimationtoolkit_b200 as pbat`` gives ``pbat.sim.vbd.Data``,
``pbat.sim.vbd.Integrator``, ``pbat.gpu.vbd.Integrator`` and the enums the reference exposes
for this path (bindings/pypbat/sim/vbd, bindings/pypbat/gpu/vbd).  Everything executes in
hand-written sm_100a CUDA kernels behind the C-ABI of ``include/vbdx.h``.
"""
from . import graph, host, meshes, sim, gpu  # noqa: F401

__all__ = ["graph", "host", "meshes", "sim", "gpu"]
