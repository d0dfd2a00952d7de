import sys, os
sys.path.insert(0,'/root/repo')
import numpy as np
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes
X, T = meshes.tet_grid(58, 58, 58, 1/58)
dbc = np.flatnonzero(X[2] == 0)
d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_chebyshev_acceleration(0.9).construct()
for kd in (0.0, 1e-4):
    vbd = pbat.gpu.vbd.Integrator(d)
    vbd.kD = kd
    ms=[]
    for s in range(40):
        vbd.step(0.01,30,1); ms.append(vbd.info["lastStepMs"])
    print("config 2 kD", kd, "ms/step", float(np.median(ms[5:])))
