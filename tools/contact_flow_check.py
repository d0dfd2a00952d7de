"""Config 3 at size (16 stacked bodies of 29^3, contact): the barrier-free sweep with its write history against the sweep
with colour barriers -- bitwise, and how long a step takes in both.   python tools/contact_flow_check.py [scale] [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
n = max(2, int(round(29 * scale)))
Xb, Tb = meshes.tet_grid(n, n, n, 1.0 / n)
X, T, B = meshes.stack_bodies(Xb, Tb, 16, axis=2, gap_frac=0.1)
for b in range(16):                       # generic lateral offsets (tests/test_gpu_contact.py: config-3 replica)
    X[0, B == b] += 0.37 * (b % 3) / n
    X[1, B == b] += 0.21 * (b % 4) / n
F = meshes.boundary_facets(T)
V = np.unique(F)
dbc = np.flatnonzero(X[2] <= X[2].min() + 0.01 * (Xb[2].max() - Xb[2].min()))
d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_surface_mesh(V, F).with_bodies(B)
     .with_dirichlet_vertices(dbc).with_contact_parameters(1e6, 0.3, 1e-3).construct())
res = {}
for mode in ("1", "0"):
    os.environ["VBDX_DATAFLOW"] = mode
    vbd = pbat.gpu.vbd.Integrator(d)
    ms, contacts = [], 0
    for s in range(steps):
        vbd.step(0.01, 20, 1)
        ms.append(vbd.info["lastStepMs"])
    _, nn, na = vbd.contact_state()
    res[mode] = dict(x=vbd.x.copy(), v=vbd.v.copy(), nn=nn.copy(), ms=float(np.median(ms[5:])), na=int(na),
                     contacts=int((nn >= 0).any(axis=1).sum()), threads=vbd.info["blockThreads"])
a, b = res["1"], res["0"]
print(json.dumps({"bodies": f"16 x {n}^3", "nV": int(X.shape[1]), "nT": int(T.shape[1]), "steps": steps,
                  "ms_per_step_barrier_free": a["ms"], "ms_per_step_barriers": b["ms"],
                  "block_threads": [a["threads"], b["threads"]], "active_vertices": a["na"], "vertices_with_contacts": a["contacts"],
                  "bitwise_equal_x": bool(np.array_equal(a["x"], b["x"])), "bitwise_equal_v": bool(np.array_equal(a["v"], b["v"])),
                  "contact_lists_equal": bool(np.array_equal(a["nn"], b["nn"])), "finite": bool(np.isfinite(a["x"]).all()),
                  "max_abs_diff_x": float(np.abs(a["x"] - b["x"]).max())}))
