"""GPU tests of the XPBD integrator against the oracle's restatement of sim/xpbd/Integrator.cpp:31-139
(its _ref build calls the reference's own ProjectBlockNeoHookean / ProjectVertexTriangle, sim/xpbd/Kernels.h) and the
reference's known-answer test."""
import numpy as np
import pytest

import oracle
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_reference_cube_doctest():
    """sim/xpbd/Integrator.cpp:213-262 (and gpu/impl/xpbd/Integrator.cu's twin): the unit cube with its own surface as collision
    mesh, one constraint per partition, dt = 1e-2, 1 iteration, 20 substeps: every vertex falls, none moves sideways."""
    P, T, F = meshes.CUBE_P, meshes.CUBE_T, meshes.CUBE_F
    d = (pbat.sim.xpbd.Data().with_volume_mesh(P, T).with_surface_mesh(np.arange(8), F)
         .with_partitions([0, 1, 2, 3, 4, 5], [0, 1, 2, 3, 4]).construct())
    xpbd = pbat.gpu.xpbd.Integrator(d)
    xpbd.step(1e-2, 1, 20)
    dx = xpbd.x.astype(np.float64) - P
    assert (dx[2] < 0).all()
    assert (np.abs(dx[:2]) < 1e-4).all()
    ref = oracle.Oracle(P, T, V=np.arange(8), F=F, B=np.zeros(8, np.int64))
    ref.xpbd_setup([0, 1, 2, 3, 4, 5], [0, 1, 2, 3, 4])
    ref.xpbd_step(1e-2, 1, 20)
    assert rel_l2(xpbd.x, ref.x) < 1e-5
    assert xpbd.info["kernelLaunches"] > 0


@pytest.mark.parametrize("clustered", [False, True])
def test_elastic_beam_against_oracle(clustered):
    """A cantilever (no collision mesh: one persistent launch per Step) with heterogeneous material, damping on the elastic
    constraints, perturbed rest shape; plain and clustered partitions.  Particle masses: 10 kg (the reference's default is
    1000 kg, sim/xpbd/Data.cpp:113-116).  With the 0.03 kg of rho V / 4 at this resolution the block Neo-Hookean constraint
    is out of fp32's reach -- its value |F| ~ 1.7 must cancel alpha~ lambda to the ~1e-7 that a particle's weight
    contributes, and fp32 positions of O(1) resolve |F| to ~1e-6 -- for ANY fp32 implementation, the reference's GPU path
    included (tools/xpbd_diag.py: displacement-relative 2-6e-2 instead of 1e-4); that case is checked for its outcome below."""
    X, T = meshes.tet_grid(10, 3, 3, 0.1)
    nV, nT = X.shape[1], T.shape[1]
    dbc = np.flatnonzero(X[0] == 0)
    rng = np.random.default_rng(1)
    X = X + 0.01 * rng.uniform(-1, 1, X.shape)
    mu, lam = pbat.sim.vbd.lame_coefficients(1e6, 0.45)
    lame = np.stack([mu * rng.uniform(0.5, 2, nT), lam * rng.uniform(0.5, 2, nT)])
    vol = np.abs(meshes.tet_volumes(X, T))
    m = np.full(nV, 10.0) * rng.uniform(0.8, 1.25, nV)
    beta = np.full(2 * nT, 1e-3)
    Pptr, Padj, GC = pbat.sim.xpbd.partition_mesh_constraints(X, T)
    data = (pbat.sim.xpbd.Data().with_volume_mesh(X, T).with_mass_inverse(1.0 / m).with_elastic_material(lame)
            .with_damping(beta, pbat.sim.xpbd.Constraint.StableNeoHookean).with_dirichlet_constrained_vertices(dbc)
            .with_partitions(Pptr, Padj))
    if clustered:
        # clusters = pairs of consecutive partitions' constraints... any grouping whose clusters of one cluster partition
        # share no vertex: here every cluster is two constraints of DIFFERENT colours that do share vertices, coloured greedily
        pairs = [[int(Padj[i]), int(Padj[i + 1])] if i + 1 < len(Padj) else [int(Padj[i])] for i in range(0, len(Padj), 2)]
        owner = np.empty(nT, np.int64)
        for c, cl in enumerate(pairs):
            owner[cl] = c
        import scipy.sparse as sp
        G = pbat.graph.mesh_adjacency_matrix(T, nV)                      # vertex x element
        S = sp.csc_matrix((np.ones(nT), (owner, np.arange(nT))), shape=(len(pairs), nT))
        CG = ((S @ G.T) @ (S @ G.T).T).tocsc()                           # clusters adjacent through a shared vertex
        SGC = pbat.graph.greedy_color(CG.indptr, CG.indices)
        SGptr, SGadj = pbat.graph.map_to_adjacency(SGC)
        Cptr = np.concatenate([[0], np.cumsum([len(c) for c in pairs])])
        Cadj = np.concatenate(pairs)
        data = data.with_cluster_partitions(SGptr, SGadj, Cptr, Cadj)
    data = data.construct()
    xpbd = pbat.sim.xpbd.Integrator(data)
    ref = oracle.Oracle(X, T, dbc=dbc, mue=lame[0], lambdae=lame[1])
    if clustered:
        # the oracle runs clusters as partitions of single constraints in the same order: cluster partition by cluster partition,
        # inside a partition the clusters' constraints -- a cluster's two constraints must run one after the other, which an
        # OpenMP loop over a partition does not guarantee: give every position in a cluster its own partition
        P2ptr, P2adj = [0], []
        for q in range(len(SGptr) - 1):
            cls = [pairs[c] for c in SGadj[SGptr[q]:SGptr[q + 1]]]
            for pos in range(2):
                members = [cl[pos] for cl in cls if len(cl) > pos]
                if members:
                    P2adj += members
                    P2ptr.append(len(P2adj))
        ref.xpbd_setup(P2ptr, P2adj, minv=1.0 / m, beta_snh=beta)
    else:
        ref.xpbd_setup(Pptr, Padj, minv=1.0 / m, beta_snh=beta)
    for _ in range(10):
        xpbd.step(0.01, 5, 4)
        ref.xpbd_step(0.01, 5, 4)
    xr = ref.x
    err, derr = rel_l2(xpbd.x, xr), np.linalg.norm(xpbd.x - xr) / np.linalg.norm(xr - X)
    print(f"XPBD beam clustered={clustered}: rel L2 = {err:.3e}, displacement-relative = {derr:.3e}, tip drop = {(xr - X)[2].min():.4f}")
    assert np.isfinite(xpbd.x).all() and err < 1e-4 and derr < 1e-3
    assert (xr - X)[2].min() < -0.01          # it did bend
    assert np.array_equal(xpbd.x[:, dbc], X[:, dbc].astype(np.float32).astype(np.float64))   # Dirichlet vertices never move
    assert np.allclose(xpbd.v, ref.v, atol=2e-3 * np.abs(ref.v).max())
    if not clustered:
        before = xpbd._impl.info["kernelLaunches"]
        xpbd._impl.step(0.01, 5, 4)
        assert xpbd._impl.info["kernelLaunches"] - before == 1   # one persistent launch per Step: 4 substeps x 5 iterations x 30+ partitions
        # light particles (rho V / 4): same outcome, fp32-limited agreement
        ml = np.bincount(T.reshape(-1), weights=np.tile(1e3 * vol / 4, 4), minlength=nV)
        dl = (pbat.sim.xpbd.Data().with_volume_mesh(X, T).with_mass_inverse(1.0 / ml).with_dirichlet_constrained_vertices(dbc)
              .with_partitions(Pptr, Padj).construct())
        light = pbat.gpu.xpbd.Integrator(dl)
        rl = oracle.Oracle(X, T, dbc=dbc)
        rl.xpbd_setup(Pptr, Padj, minv=1.0 / ml)
        for _ in range(10):
            light.step(0.01, 5, 4)
            rl.xpbd_step(0.01, 5, 4)
        dg, dr = (light.x - X)[2].min(), (rl.x - X)[2].min()
        assert np.isfinite(light.x).all() and dr < -0.03 and abs(dg - dr) < 0.15 * abs(dr), (dg, dr)
        assert np.linalg.norm(light.x - rl.x) / np.linalg.norm(rl.x - X) < 0.25


def test_contact_against_oracle():
    """A small body dropped on a fixed-base body (generic lateral offset, see tests/test_gpu_contact.py): contact lists and
    trajectory against the oracle (same detector semantics: gpu/impl/contact/VertexTriangleMixedCcdDcd.cu)."""
    n = 4
    Xb, Tb = meshes.tet_grid(n, n, n, 0.5 / n)
    Xt = Xb * 0.6 + np.array([[0.11], [0.07], [0.58]])
    X = np.concatenate([Xb, Xt], axis=1)
    T = np.concatenate([Tb, Tb + Xb.shape[1]], axis=1)
    B = np.concatenate([np.zeros(Xb.shape[1], np.int64), np.ones(Xb.shape[1], np.int64)])
    F = meshes.boundary_facets(T)
    V = np.unique(F)
    dbc = np.flatnonzero(X[2] == 0)
    v = np.zeros_like(X)
    v[2, Xb.shape[1]:] = -0.5
    vol = np.abs(meshes.tet_volumes(X, T))
    m = np.full(X.shape[1], 10.0)
    Pptr, Padj, _ = pbat.sim.xpbd.partition_mesh_constraints(X, T)
    muV = np.full(V.size, 1.0)
    data = (pbat.sim.xpbd.Data().with_volume_mesh(X, T).with_surface_mesh(V, F).with_bodies(B).with_velocity(v)
            .with_mass_inverse(1.0 / m).with_collision_penalties(muV).with_friction_coefficients(0.4, 0.3)
            .with_dirichlet_constrained_vertices(dbc).with_partitions(Pptr, Padj).construct())
    xpbd = pbat.gpu.xpbd.Integrator(data)
    ref = oracle.Oracle(X, T, v=v, dbc=dbc, B=B, V=V, F=F)
    ref.xpbd_setup(Pptr, Padj, minv=1.0 / m, muV=muV, muS=0.4, muD=0.3)
    touched, agree_until = 0, None
    for s in range(30):
        xpbd.step(0.01, 4, 5)
        ref.xpbd_step(0.01, 4, 5)
        act, nn, _ = xpbd.contact_state()
        touched = max(touched, int((nn >= 0).any(axis=1).sum()))
        same = np.array_equal(np.sort(nn, axis=1), np.sort(ref.get("nn").reshape(-1, 8), axis=1))
        if not same and agree_until is None:
            agree_until = s
        if agree_until is None:
            # (same nearest triangles; whether a pair is projected at all -- signed distance <= 0, projection inside the
            # triangle -- is still decided in fp32 here and in double there)
            assert rel_l2(xpbd.x, ref.x) < 1e-3, (s, rel_l2(xpbd.x, ref.x))
    err = rel_l2(xpbd.x, ref.x)
    print(f"XPBD contact: vertices in contact (max) = {touched}, contact lists equal until step {agree_until}, rel L2 after 30 steps = {err:.3e}")
    assert touched > 5, "the bodies never touched"
    assert agree_until is None or agree_until >= 12
    assert np.isfinite(xpbd.x).all() and err < 2e-2
    assert xpbd.x[2, B == 1].min() > 0.4   # resting on the lower body, not through it
