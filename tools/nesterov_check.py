"""Scratch: per-step parity of the device Nesterov solve against the oracle, next to what rounding the oracle's iterates to
fp32 after every operation does on its own (CPU emulation)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

X, T = meshes.tet_grid(12, 4, 4, 0.05)
dbc = np.flatnonzero(X[0] == 0)
f32 = lambda a: a.astype(np.float32).astype(np.float64)


def nesterov_py(o, dt, iters, L, start, substeps):
    sdt = dt / substeps
    for _ in range(substeps):
        o.step(sdt, 0, 1)
        xt = o.get("xt")
        x = f32(o.x); o.x = x
        xkm1 = x.copy(); alpha = 1.0 / L; lam = 0.0; beta = 0.0
        for k in range(iters):
            on = start < k
            if on:
                yk = f32(x + np.float32(beta) * (x - xkm1))
            o.sweeps(sdt, 1); x = f32(o.x); o.x = x
            if on:
                x = f32(yk - np.float32(alpha) * (x - xkm1)); o.x = x
                lk = lam; lam = (1 + np.sqrt(1 + 4 * lam * lam)) / 2; beta = (lk - 1) / lam
        o.v = f32((x - xt) / sdt)
    return x


for L, start, substeps in ((4.0, 3, 1), (10.0, 0, 2), (1.0, 3, 1)):
    d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_nesterov_acceleration(L, start).construct()
    vbd = pbat.sim.vbd.Integrator(d)
    ref = oracle.Oracle(X, T, dbc=dbc, colors=d.colors); ref.set_acceleration(oracle.ACCEL_NESTEROV, L=L, start=start)
    emu = oracle.Oracle(X, T, dbc=dbc, colors=d.colors)
    for s in range(3):
        x0, v0 = ref.x, ref.v
        vbd.x, vbd.v = x0, v0
        emu.x, emu.v = x0, v0
        vbd.step(0.01, 10, substeps); ref.step(0.01, 10, substeps)
        xe = nesterov_py(emu, 0.01, 10, L, start, substeps)
        upd = np.linalg.norm(ref.x - x0)
        print(f"L={L} start={start} substeps={substeps} step {s}: update {upd:.3e}  device-vs-oracle {np.linalg.norm(vbd.x-ref.x)/upd:.3e}  "
              f"fp32-emulation-vs-oracle {np.linalg.norm(xe-ref.x)/upd:.3e}  device-vs-emulation {np.linalg.norm(vbd.x-xe)/upd:.3e}", flush=True)
