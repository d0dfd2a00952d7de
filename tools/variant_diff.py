import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes
X, T = meshes.tet_grid(12, 6, 5, 0.1)
dbc = np.flatnonzero(X[0] == 0)
for kD in (0.0, 1e-4):
    for steps, iters, sub in ((1, 1, 1), (1, 3, 1), (1, 7, 2), (3, 7, 2)):
        out = []
        for variant in (1, 2, 3):
            d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_chebyshev_acceleration(0.9).with_rayleigh_damping(kD).construct()
            vbd = pbat.gpu.vbd.Integrator(d, kernel_variant=variant, tile_iters=2)
            for _ in range(steps):
                vbd.step(0.01, iters, sub)
            out.append(vbd.x)
        print(f"kD={kD} steps={steps} iters={iters} sub={sub}: |v1-v2|={np.abs(out[0]-out[1]).max():.3e} |v1-v3|={np.abs(out[0]-out[2]).max():.3e} nnz13={(out[0]!=out[2]).sum()}")
