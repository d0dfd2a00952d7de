"""BASELINE configs[2] (16 stacked bodies, contact): device ms of every step, barrier-free sweep vs colour barriers, and the number
of active vertices every 10 steps -- shows the two phases (free fall, then contact) and that both sweeps take the same
contact decisions.   python tools/config3_phases.py"""
import sys, time, os
sys.path.insert(0,'/root/repo')
import numpy as np
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes
n=29
Xb, Tb = meshes.tet_grid(n, n, n, 1.0 / n)
X, T, B = meshes.stack_bodies(Xb, Tb, 16, axis=2, gap_frac=0.1)
F = meshes.boundary_facets(T); V = np.unique(F)
dbc = np.flatnonzero(X[2] <= X[2].min() + 0.01)
d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_surface_mesh(V, F).with_bodies(B)
     .with_dirichlet_vertices(dbc).with_contact_parameters(1e6, 0.3, 1e-3).construct())
for mode in ("1","0"):
    os.environ["VBDX_DATAFLOW"]=mode
    vbd = pbat.gpu.vbd.Integrator(d)
    ms=[]; na=[]
    for s in range(70):
        vbd.step(0.01,20,1); ms.append(vbd.info["lastStepMs"])
        if s%10==9: na.append(int(vbd.contact_state()[2]))
    print(mode, " ".join(f"{m:.2f}" for m in ms)); print(na)
