"""Small-mesh latency: BASELINE config 1 (25x9x9 cantilever, 10,125 tets, 20 iterations) and the reference's doctest
cube under every kernel variant; prints ms/step (device-timed) and checks the variants agree bitwise.
  python tools/small_mesh_variants.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

NAMES = {0: "default", 1: "direct", 3: "pipelined", 4: "cluster"}
out = {}
for label, (X, T, dbc) in {
    "config1": (*meshes.tet_grid(25, 9, 9, 0.04), None),
    "beam_6x3x3": (*meshes.tet_grid(6, 3, 3, 0.1), None),
    "grid_16^3": (*meshes.tet_grid(16, 16, 16, 1 / 16), None),
    "grid_24^3": (*meshes.tet_grid(24, 24, 24, 1 / 24), None),
}.items():
    dbc = np.flatnonzero(X[0] == 0)
    for cheb in (None, 0.9):
        d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc)
        if cheb:
            d = d.with_chebyshev_acceleration(cheb)
        d = d.construct()
        xs, row = [], {}
        for variant in (0, 1, 3, 4):
            try:
                vbd = pbat.gpu.vbd.Integrator(d, kernel_variant=variant)
            except Exception as e:  # noqa: BLE001
                row[NAMES[variant]] = str(e)[:60]
                continue
            ms = []
            for _ in range(60):
                vbd.step(0.01, 20, 1)
                ms.append(vbd.info["lastStepMs"])
            row[NAMES[variant]] = round(float(np.median(ms[10:])), 4)
            xs.append(vbd.x.copy())
        row["bitwise_equal"] = bool(all(np.array_equal(xs[0], x) for x in xs[1:]))
        row["tets"] = int(T.shape[1])
        out[f"{label}{'_cheb' if cheb else ''}"] = row
print(json.dumps(out, indent=1))
