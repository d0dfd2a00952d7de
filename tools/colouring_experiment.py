"""How much the number of colours is worth on configs[1] ("fewer colours = fewer barriers" -- here: fewer
dependent tile rounds per sweep).  The 5-tet-per-cube grid has a balanced 5-colouring: vertices of odd parity i+j+k are
pairwise non-adjacent (face diagonals only join even vertices) = one colour with 20 % of the incidences; the even ones form
an FCC lattice, 4-coloured by (i, j, k) mod 2.   python tools/colouring_experiment.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 58
X, T = meshes.tet_grid(n, n, n, 1 / n)
nV = X.shape[1]
ids = np.arange(nV)
k = ids % (n + 1); j = (ids // (n + 1)) % (n + 1); i = ids // ((n + 1) * (n + 1))
odd = ((i + j + k) & 1) == 1
cls = {(0, 0, 0): 0, (1, 1, 0): 1, (1, 0, 1): 2, (0, 1, 1): 3}
five = np.where(odd, 4, [cls.get((a & 1, b & 1, c & 1), -1) for a, b, c in zip(i, j, k)])
assert (five >= 0).all()
for a in range(4):
    for b in range(a + 1, 4):
        assert (five[T[a]] != five[T[b]]).all()          # a proper colouring: no tet has two vertices of one colour
dbc = np.flatnonzero(X[2] == 0)
x0 = X + 0.05 / n * np.random.default_rng(0).uniform(-1, 1, X.shape)
x0[:, dbc] = X[:, dbc]
out = []
first = pbat.graph.mesh_greedy_color(T, nV, ordering=2, selection=1)     # LargestDegree / FirstAvailable: the reference's other selection
for name, colors in (("reference default (LargestDegree / LeastUsed)", None), ("balanced 5-colouring of the grid", five),
                     ("LargestDegree / FirstAvailable", first)):
    for tile_iters in ((0,) if colors is None else (0, 10, 11, 12, 13, 14, 16)):
        d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_chebyshev_acceleration(0.9).construct()
        if colors is not None:
            d.colors = colors.astype(np.int64)
        vbd = pbat.gpu.vbd.Integrator(d, tile_iters=tile_iters)
        vbd.x = x0.astype(np.float32)
        ms = []
        for s in range(30):
            vbd.step(0.01, 30, 1)
            ms.append(vbd.info["lastStepMs"])
        info = vbd.info
        rec = {"colouring": name, "colours": int(info["nColors"]), "tile_iters": tile_iters, "tiles": int(info["nTiles"]),
               "ms_per_step": float(np.median(ms[5:])), "vertex_iterations_per_s": info["nActiveVertices"] * 30 / (np.median(ms[5:]) * 1e-3)}
        if n <= 20:                                       # parity: the oracle sweeps whatever colouring it is given
            ref = oracle.Oracle(X, T, dbc=dbc, colors=d.colors, accel=oracle.ACCEL_CHEBYSHEV, rho=0.9)
            ref.x = x0.astype(np.float32).astype(np.float64)
            for s in range(30):
                ref.step(0.01, 30, 1)
            rec["rel_l2_vs_oracle"] = float(np.linalg.norm(vbd.x - ref.x) / np.linalg.norm(ref.x))
        out.append(rec)
        print(json.dumps(rec), flush=True)
