"""Soak test of the contact write history: a stack of bodies that fall, collide, slide and settle, stepped
thousands of times with the barrier-free sweep and with colour barriers (VBDX_DATAFLOW=0); positions, velocities and contact
lists compared BITWISE every `every` steps.    python tools/soak_contact.py [bodies] [n] [steps] [every]"""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
every = int(sys.argv[4]) if len(sys.argv) > 4 else 500
Xb, Tb = meshes.tet_grid(n, n, n, 1.0 / n)
X, T, B = meshes.stack_bodies(Xb, Tb, nb, axis=2, gap_frac=0.1)
for b in range(nb):
    X[0, B == b] += 0.37 * (b % 3) / n
    X[1, B == b] += 0.21 * (b % 4) / n
F = meshes.boundary_facets(T)
V = np.unique(F)
dbc = np.flatnonzero(X[2] <= X[2].min() + 0.01)
d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_surface_mesh(V, F).with_bodies(B)
     .with_dirichlet_vertices(dbc).with_contact_parameters(1e6, 0.3, 1e-3).with_chebyshev_acceleration(0.8).construct())
os.environ["VBDX_DATAFLOW"] = "1"
flow = pbat.gpu.vbd.Integrator(d, kernel_variant=3)
os.environ["VBDX_DATAFLOW"] = "0"
bar = pbat.gpu.vbd.Integrator(d, kernel_variant=3)
os.environ.pop("VBDX_DATAFLOW")
assert flow.info["blockThreads"] != bar.info["blockThreads"]
t = time.time()
ok = True
peak = 0
for s in range(1, steps + 1):
    flow.step(0.01, 10, 1 + (s % 3 == 0))          # every third step with two substeps
    bar.step(0.01, 10, 1 + (s % 3 == 0))
    if s % every == 0 or s == steps:
        xf, xb, vf, vb = flow.x, bar.x, flow.v, bar.v
        cf, cb = flow.contact_state(), bar.contact_state()
        same = np.array_equal(xf, xb) and np.array_equal(vf, vb) and np.array_equal(cf[1], cb[1]) and np.array_equal(cf[0], cb[0])
        ok &= same
        peak = max(peak, int((cf[1] >= 0).any(axis=1).sum()))
        print(f"step {s}: bitwise equal = {same}  sha1(x) = {hashlib.sha1(xf.tobytes()).hexdigest()[:12]}  finite = {bool(np.isfinite(xf).all())}  "
              f"active = {int(cf[2])}  vertices in contact = {int((cf[1] >= 0).any(axis=1).sum())}  ({time.time()-t:.1f} s)", flush=True)
print("CONTACT SOAK", "PASSED" if ok and peak > 0 else "FAILED", f"{steps} steps, {nb} bodies of {n}^3 cubes ({T.shape[1]} tets), Chebyshev, 10 iterations")
sys.exit(0 if ok and peak > 0 else 1)
