"""Scratch: which ingredient separates the device XPBD solve from the oracle on the cantilever test."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

X0, T = meshes.tet_grid(10, 3, 3, 0.1)
nV, nT = X0.shape[1], T.shape[1]
dbc = np.flatnonzero(X0[0] == 0)
mu, lam = pbat.sim.vbd.lame_coefficients(1e6, 0.45)
for name, hetero, damp, perturb, realmass, iters, sub in (("light particles (rho V / 4)", 0, 0, 0, 1, 5, 4), ("default mass", 0, 0, 0, 0, 5, 4),
                                                           ("default mass, hetero + damped + perturbed", 1, 1, 1, 0, 5, 4),
                                                           ("default mass, all, 10 its 1 substep", 1, 1, 1, 0, 10, 1), ("mass 10", 0, 0, 0, 2, 5, 4)):
    rng = np.random.default_rng(1)
    X = X0 + (0.01 * rng.uniform(-1, 1, X0.shape) if perturb else 0)
    lame = np.stack([mu * (rng.uniform(0.5, 2, nT) if hetero else np.ones(nT)), lam * (rng.uniform(0.5, 2, nT) if hetero else np.ones(nT))])
    vol = np.abs(meshes.tet_volumes(X, T))
    m = np.bincount(T.reshape(-1), weights=np.tile(1e3 * vol / 4, 4), minlength=nV) if realmass == 1 else np.full(nV, 1e3 if realmass == 0 else 10.0)
    beta = np.full(2 * nT, 1e-3 if damp else 0.0)
    Pptr, Padj, GC = pbat.sim.xpbd.partition_mesh_constraints(X, T)
    data = (pbat.sim.xpbd.Data().with_volume_mesh(X, T).with_mass_inverse(1.0 / m).with_elastic_material(lame)
            .with_damping(beta, 0).with_dirichlet_constrained_vertices(dbc).with_partitions(Pptr, Padj).construct())
    xpbd = pbat.sim.xpbd.Integrator(data)
    ref = oracle.Oracle(X, T, dbc=dbc, mue=lame[0], lambdae=lame[1])
    ref.xpbd_setup(Pptr, Padj, minv=1.0 / m, beta_snh=beta)
    out = []
    for s in range(10):
        xpbd.step(0.01, iters, sub)
        ref.xpbd_step(0.01, iters, sub)
        out.append(np.linalg.norm(xpbd.x - ref.x) / np.linalg.norm(ref.x - X))
    print(f"{name:44s} displacement-relative error after steps 1, 2, 5, 10: {out[0]:.2e} {out[1]:.2e} {out[4]:.2e} {out[9]:.2e}   tip drop {(ref.x - X)[2].min():.4f}", flush=True)
