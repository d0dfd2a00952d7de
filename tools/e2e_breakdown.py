"""Where the end-to-end step (host buffers in, host buffers out) spends its time on configs[1]:
python tools/e2e_breakdown.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

X, T = meshes.tet_grid(58, 58, 58, 1 / 58)
dbc = np.flatnonzero(X[2] == 0)
d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_chebyshev_acceleration(0.9).construct()
vbd = pbat.gpu.vbd.Integrator(d)
nV = X.shape[1]
xin = pbat.host.pinned_empty((3, nV), np.float32)
xout = pbat.host.pinned_empty((3, nV), np.float32)
xin[...] = X
t = np.zeros(4)
n = 200
for it in range(n + 10):
    if it == 10:
        t[:] = 0
    a = time.perf_counter()
    vbd.x = xin
    b = time.perf_counter()
    vbd.step(0.01, 30, 1)
    c = time.perf_counter()
    vbd.positions(out=xout)
    e = time.perf_counter()
    xin, xout = xout, xin
    t += [b - a, c - b, e - c, vbd.info["lastStepMs"] * 1e-3]
print("per step: set x %.1f us, step %.1f us (device %.1f us), get x %.1f us, total %.1f us" % tuple(
    1e6 * v / n for v in (t[0], t[1], t[3], t[2], t[0] + t[1] + t[2])))
