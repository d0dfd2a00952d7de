"""Strong scaling of BASELINE config 4 (SURVEY.md 8d "C4"): ONE 117^3-cube tet block (8,008,065 tets, 1,643,032
vertices), Stable Neo-Hookean, 30 iterations/step, Chebyshev rho = 0.9, z = 0 face fixed, domain-decomposed into
x-slabs over the GPUs of the box (N = 1: the plain single-GPU integrator).

  python tools/strong_scaling.py [grid] [steps]                                  # N = 1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/strong_scaling.py

Prints one JSON line on rank 0: vertex-iterations/s of the whole job (device time, max over ranks), ms/step and the
HBM roofline fraction of the record format's bytes per GPU.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes
from bench import format_bytes_per_vertex_iteration

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 117
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ITERS, RHO, DT, WARMUP = 30, 0.9, 0.01, 2

t_setup = time.perf_counter()
X, T = meshes.tet_grid(grid, grid, grid, 1.0 / grid)
nV = X.shape[1]
dbc = np.flatnonzero(X[2] == 0)
x0 = X + 0.05 / grid * np.random.default_rng(0).uniform(-1, 1, X.shape)
x0[:, dbc] = X[:, dbc]
dist = None
if world > 1:
    import torch.distributed as dist
    from physicsbasedanimationtoolkit_b200.dist import DomainDecomposedIntegrator

    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    dd = DomainDecomposedIntegrator(X, T, dbc=dbc, rho_chebyshev=RHO, axis=0)
    vbd, lp = dd.vbd, dd.local
    vbd.x = np.ascontiguousarray(x0[:, lp.l2g], dtype=np.float32)
else:
    data = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_chebyshev_acceleration(RHO).construct()
    vbd = pbat.gpu.vbd.Integrator(data)
    vbd.x = np.ascontiguousarray(x0, dtype=np.float32)
t_setup = time.perf_counter() - t_setup


def barrier():
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()


barrier()
for _ in range(WARMUP):
    vbd.step(DT, ITERS, 1)
barrier()
stream = torch.cuda.Stream()
vbd.use_stream(stream.cuda_stream)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(stream)
for _ in range(steps):
    vbd.step_async(DT, ITERS, 1)
ev1.record(stream)
vbd.synchronize()
barrier()
ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
ok = torch.tensor([float(np.isfinite(vbd.x).all())], device="cuda")
if dist is not None:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
ms = float(ms[0])
if rank == 0:
    active = np.ones(nV, bool)
    active[dbc] = False
    B = format_bytes_per_vertex_iteration(vbd.info)["total"]  # bytes the record format moves per vertex solve (this rank's counts)
    n_active = int(active.sum())
    value = n_active * ITERS * steps / (ms * 1e-3)
    peak = 6552.0
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    print(json.dumps({"workload": f"configs[3]: {grid}^3-cube block, {T.shape[1]} tets, {nV} vertices, {ITERS} iterations/step, Chebyshev {RHO}",
                      "n_gpus": world, "scaling": "strong", "steps": steps, "warmup": WARMUP, "ms_per_step": ms / steps,
                      "value": value, "unit": "vertex-iterations/s", "steps_per_s": steps / (ms * 1e-3),
                      "roofline_frac_per_gpu": value * B / 1e9 / world / peak, "bytes_per_vertex_iteration": B,
                      "finite": bool(ok[0] > 0), "setup_s": t_setup}), flush=True)
if dist is not None:
    dist.destroy_process_group()
