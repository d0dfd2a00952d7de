"""Soak test of the barrier-free sweep's tag protocol: many steps of one scene with the barrier-free kernel
and with colour barriers (VBDX_DATAFLOW=0), positions and velocities compared BITWISE every `every` steps.
    python tools/soak.py [grid] [steps] [every]"""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 58
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
every = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
X, T = meshes.tet_grid(n, n, n, 1.0 / n)
dbc = np.flatnonzero(X[2] == 0)
data = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_chebyshev_acceleration(0.9).construct()
x0 = (X + 0.05 / n * np.random.default_rng(0).uniform(-1, 1, X.shape)).astype(np.float32)
x0[:, dbc] = X[:, dbc]
os.environ["VBDX_DATAFLOW"] = "1"
flow = pbat.gpu.vbd.Integrator(data)
os.environ["VBDX_DATAFLOW"] = "0"
bar = pbat.gpu.vbd.Integrator(data)
os.environ.pop("VBDX_DATAFLOW")
flow.x = x0
bar.x = x0
t = time.time()
ok = True
for s in range(1, steps + 1):
    flow.step_async(0.01, 30, 1)
    bar.step_async(0.01, 30, 1)
    if s % every == 0 or s == steps:
        flow.synchronize(), bar.synchronize()
        xf, xb, vf, vb = flow.x, bar.x, flow.v, bar.v
        same = np.array_equal(xf, xb) and np.array_equal(vf, vb)
        ok &= same
        print(f"step {s}: bitwise equal = {same}  sha1(x) = {hashlib.sha1(xf.tobytes()).hexdigest()[:12]}  finite = {bool(np.isfinite(xf).all())} "
              f"non-finite sentinel = {flow.info['nonFiniteVertices']}  ({time.time()-t:.1f} s)", flush=True)
print("SOAK", "PASSED" if ok else "FAILED", f"{steps} steps of {T.shape[1]} tets, 30 iterations each")
sys.exit(0 if ok else 1)
