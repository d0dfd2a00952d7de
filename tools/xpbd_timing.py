"""XPBD at size: the 16-body stack of BASELINE configs[2] (1.95 M tets, contact) and a 58^3 block under
gravity, per step of dt = 0.01 with 10 substeps x 1 iteration (the substep-heavy regime XPBD is used in).
python tools/xpbd_timing.py [scale]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
for name in ("block", "stack"):
    t0 = time.time()
    if name == "block":
        n = max(2, int(round(58 * scale)))
        X, T = meshes.tet_grid(n, n, n, 1.0 / n)
        dbc = np.flatnonzero(X[2] == 0)
        Pptr, Padj, _ = pbat.sim.xpbd.partition_mesh_constraints(X, T)
        d = (pbat.sim.xpbd.Data().with_volume_mesh(X, T).with_mass_inverse(np.full(X.shape[1], 0.1))
             .with_dirichlet_constrained_vertices(dbc).with_partitions(Pptr, Padj).construct())
    else:
        n = max(2, int(round(29 * scale)))
        Xb, Tb = meshes.tet_grid(n, n, n, 1.0 / n)
        X, T, B = meshes.stack_bodies(Xb, Tb, 16, axis=2, gap_frac=0.1)
        F = meshes.boundary_facets(T)
        V = np.unique(F)
        dbc = np.flatnonzero(X[2] <= X[2].min() + 0.01)
        Pptr, Padj, _ = pbat.sim.xpbd.partition_mesh_constraints(X, T)
        d = (pbat.sim.xpbd.Data().with_volume_mesh(X, T).with_surface_mesh(V, F).with_bodies(B)
             .with_mass_inverse(np.full(X.shape[1], 0.1)).with_collision_penalties(np.full(V.size, 1.0))
             .with_dirichlet_constrained_vertices(dbc).with_partitions(Pptr, Padj).construct())
    t1 = time.time()
    xp = pbat.gpu.xpbd.Integrator(d)
    t2 = time.time()
    steps, iters, substeps = 40, 1, 10
    for _ in range(3):
        xp.step(0.01, iters, substeps)
    ms = []
    for _ in range(steps):
        xp.step(0.01, iters, substeps)
        ms.append(xp.info["lastStepMs"] if isinstance(xp.info, dict) and "lastStepMs" in xp.info else float("nan"))
    info = xp.info
    print(json.dumps({"scene": name, "nV": int(X.shape[1]), "nT": int(T.shape[1]), "partitions": int(len(Pptr) - 1), "substeps": substeps,
                      "iterations": iters, "ms_per_step_median": float(np.median(ms)), "ms_first_10": float(np.median(ms[:10])),
                      "ms_last_10": float(np.median(ms[-10:])),
                      "constraint_projections_per_s": T.shape[1] * iters * substeps / (np.median(ms) * 1e-3),
                      "host_construct_s": round(t1 - t0, 2), "create_s": round(t2 - t1, 2), "finite": bool(np.isfinite(xp.x).all()),
                      "info": {k: (int(v) if isinstance(v, (int, np.integer)) else v) for k, v in (info.items() if isinstance(info, dict) else [])}}))
