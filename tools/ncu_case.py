"""One config-2 (or other grid) integrator stepping a few times: the target of ncu captures.  python tools/ncu_case.py [grid] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 58
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
X, T = meshes.tet_grid(n, n, n, 1.0 / n)
dbc = np.flatnonzero(X[2] == 0)
data = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_chebyshev_acceleration(0.9).construct()
vbd = pbat.gpu.vbd.Integrator(data)
xp = (X + 0.05 / n * np.random.default_rng(0).uniform(-1, 1, X.shape)).astype(np.float32)
xp[:, dbc] = X[:, dbc]
vbd.x = xp
for _ in range(steps):
    vbd.step(0.01, 30, 1)
print("ms/step", vbd.info["lastStepMs"], "finite", bool(np.isfinite(vbd.x).all()))
