"""Scratch: per-step divergence of the contact path from the contact oracle on a scaled-down configs[2] stack."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
muC = float(sys.argv[2]) if len(sys.argv) > 2 else 1e6
Xb, Tb = meshes.tet_grid(n, n, n, 1.0 / n)
X, T, B = meshes.stack_bodies(Xb, Tb, 3, axis=2, gap_frac=0.1)
shift = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
for b in range(3):  # generic position of every body over the one below (vertex-on-vertex projections are degenerate)
    X[0, B == b] += shift * 0.37 * b / n
    X[1, B == b] += shift * 0.21 * b / n
F = meshes.boundary_facets(T)
V = np.unique(F)
dbc = np.flatnonzero(X[2] <= X[2].min() + 0.01)
d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_surface_mesh(V, F).with_bodies(B)
     .with_dirichlet_vertices(dbc).with_contact_parameters(muC, 0.3, 1e-3).construct())
vbd = pbat.gpu.vbd.Integrator(d)
ref = oracle.Oracle(X, T, dbc=dbc, colors=d.colors, B=B, V=V, F=F, muC=muC, muF=0.3, epsv=1e-3)
for s in range(50):
    vbd.step(0.01, 20, 1)
    ref.step(0.01, 20, 1)
    act, nn, na = vbd.contact_state()
    rnn = ref.get("nn").reshape(-1, 8) if ref.get("nn").size else None
    ract = ref.get("active")
    same_nn = None if rnn is None else bool(np.array_equal(np.sort(nn, axis=1), np.sort(rnn, axis=1)))
    err = np.linalg.norm(vbd.x - ref.x) / np.linalg.norm(ref.x)
    print(f"step {s}: rel L2 {err:.3e}  gpu contacts {(nn>=0).any(axis=1).sum()} oracle {(rnn>=0).any(axis=1).sum() if rnn is not None else -1} "
          f"nn equal {same_nn} active equal {bool(np.array_equal(act, ract.astype(bool)))} nActive gpu {na}", flush=True)
