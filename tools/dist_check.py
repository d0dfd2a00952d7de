"""Multi-GPU domain decomposition check (run under torchrun): parity against the single-GPU integrator and
timing.  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py [n] [steps]"""
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
iters = 20
cases = [  # (chebyshev rho, Dirichlet set, substeps, reset positions every step)
    (None, "x0", 1, False),
    (0.9, "x0", 1, False),
    (0.9, "z0", 2, True),    # constrained face crossing the interfaces, substeps, host writes between steps
    (None, "z0", 1, True),
]
for cheb, fixed, substeps, poke in cases:
    X, T = meshes.tet_grid(n * world, n, n, 1.0 / n)          # a beam: one n^3 chunk per GPU (weak scaling shape)
    dbc = np.flatnonzero(X[0] == 0) if fixed == "x0" else np.flatnonzero(X[2] == 0)
    colors = pbat.graph.mesh_greedy_color(T, X.shape[1])
    dd = DomainDecomposedIntegrator(X, T, dbc=dbc, rho_chebyshev=cheb, colors=colors, axis=0)

    def advance(integ, count, local=None):
        for _ in range(count):
            if poke:  # the caller rewrites its state between steps (owned part only matters)
                x = integ.x
                integ.x = x
            integ.step(0.01, iters, substeps)

    advance(dd.vbd, 2)
    torch.cuda.synchronize(); dist.barrier()
    t = time.perf_counter()
    advance(dd.vbd, steps)
    torch.cuda.synchronize(); dist.barrier()
    el = (time.perf_counter() - t) / steps
    xg = dd.gather_x()
    ms = dd.vbd.info["lastStepMs"]
    if rank == 0:
        d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc)
        if cheb:
            d = d.with_chebyshev_acceleration(cheb)
        d = d.construct()
        assert np.array_equal(d.colors, colors)
        ref = pbat.gpu.vbd.Integrator(d)
        advance(ref, steps + 2)
        xr = ref.x
        err = np.linalg.norm(xg - xr) / np.linalg.norm(xr)
        disp = np.linalg.norm(xg - xr) / np.linalg.norm(xr - X)
        nact = X.shape[1] - dbc.size
        print(f"world={world} n={n} cheb={cheb} fixed={fixed} substeps={substeps} poke={poke}: nV={X.shape[1]} nT={T.shape[1]} "
              f"send entries rank0={dd.n_send} "
              f"rel L2 vs single GPU = {err:.3e} (displacement-relative {disp:.3e}) max|dx|={np.abs(xg-xr).max():.3e}; "
              f"step {el*1e3:.3f} ms wall, {ms:.3f} ms device (rank 0) -> {nact*iters*substeps/el/1e9:.3f} Gvert-it/s; single-GPU step {ref.info['lastStepMs']:.3f} ms", flush=True)
    dist.barrier()
    del dd
dist.destroy_process_group()
