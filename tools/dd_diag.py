"""Scratch diagnostics of the domain decomposition at strong-scaling sizes: per-rank step time and halo statistics, next to
the same half block alone on one GPU.  torchrun ... tools/dd_diag.py [grid]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes
from physicsbasedanimationtoolkit_b200.dist import DomainDecomposedIntegrator

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 117
X, T = meshes.tet_grid(grid, grid, grid, 1.0 / grid)
dbc = np.flatnonzero(X[2] == 0)
x0 = X + 0.05 / grid * np.random.default_rng(0).uniform(-1, 1, X.shape)
x0[:, dbc] = X[:, dbc]
dd = DomainDecomposedIntegrator(X, T, dbc=dbc, rho_chebyshev=0.9, axis=0)
vbd, lp = dd.vbd, dd.local
vbd.x = np.ascontiguousarray(x0[:, lp.l2g], dtype=np.float32)
info = vbd.info
for _ in range(3):
    vbd.step(0.01, 30, 1)
vbd.dist_stats()
torch.cuda.synchronize(); dist.barrier()
ms = []
for _ in range(8):
    vbd.step(0.01, 30, 1)
    ms.append(vbd.info["lastStepMs"])
st = vbd.dist_stats()
print(f"rank {rank}: nV={info['nV']} ghosts={info['nGhosts']} tiles={info['nTiles']} grid={info['gridBlocks']}x{info['blockThreads']} "
      f"step ms {np.min(ms):.3f}/{np.median(ms):.3f}  halo per step: late {st['late_ghosts']/8:.0f} epoch_waits {st['epoch_waits']/8:.1f} wait_us {st['epoch_wait_ns']/8e3:.1f}", flush=True)
dist.barrier()
# the same slab alone (ghost layer turned into constrained vertices): what the rank could do without communication
d = pbat.sim.vbd.Data().with_volume_mesh(lp.X, lp.T).with_dirichlet_vertices(np.concatenate([lp.dbc, lp.ghost_local])).with_chebyshev_acceleration(0.9).construct()
d.colors = lp.colors
alone = pbat.gpu.vbd.Integrator(d, n_colors=int(dd.colors.max()) + 1)
alone.x = np.ascontiguousarray(x0[:, lp.l2g], dtype=np.float32)
for _ in range(3):
    alone.step(0.01, 30, 1)
ms = []
for _ in range(8):
    alone.step(0.01, 30, 1)
    ms.append(alone.info["lastStepMs"])
print(f"rank {rank}: the same slab alone (ghosts frozen): step ms {np.min(ms):.3f}/{np.median(ms):.3f} tiles={alone.info['nTiles']}", flush=True)
dist.barrier()
dist.destroy_process_group()
