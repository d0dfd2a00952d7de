"""Tiny cases for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool memcheck python tools/sanitize_case.py
Every kernel family runs once: set-up, the step kernel in its direct / pipelined / one-cluster forms, Chebyshev,
St. Venant-Kirchhoff with the guarded step, Anderson, the contact prologue (radix sort, LBVH, active set), the
stand-alone BVH queries, objective evaluation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

X, T = meshes.tet_grid(5, 3, 3, 0.1)
dbc = np.flatnonzero(X[0] == 0)
for variant in (1, 3, 4):
    d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_chebyshev_acceleration(0.8).construct()
    vbd = pbat.gpu.vbd.Integrator(d, kernel_variant=variant)
    vbd.step(0.01, 3, 2)
    assert np.isfinite(vbd.x).all()
d = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc)
     .with_hyper_elastic_energy(pbat.sim.vbd.HyperElasticEnergy.SaintVenantKirchhoff).construct())
vbd = pbat.sim.vbd.Integrator(d)
vbd.line_search_guard = True
vbd.step(0.01, 3, 1)
vbd.objective_function_gradient(vbd.x, vbd.x, 0.01)
d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_anderson_acceleration(3).construct()
pbat.gpu.vbd.Integrator(d).step(0.01, 5, 1)
# the lean barrier-free kernel (variant 3 default) with Rayleigh damping, then the pipelined kernel's own barrier-free sweep
d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_chebyshev_acceleration(0.8).construct()
vbd = pbat.gpu.vbd.Integrator(d, kernel_variant=3)
vbd.kD = 0.01
vbd.step(0.01, 4, 2)
os.environ["VBDX_FLOW"] = "0"
vbd = pbat.gpu.vbd.Integrator(d, kernel_variant=3)
vbd.step(0.01, 4, 2)
del os.environ["VBDX_FLOW"]
assert np.isfinite(vbd.x).all()
d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_nesterov_acceleration(1.0, 1).construct()
pbat.gpu.vbd.Integrator(d).step(0.01, 4, 1)
d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_trust_region_acceleration(0.2, 2.0, False).construct()
pbat.gpu.vbd.Integrator(d).step(0.01, 6, 1)
# XPBD: partitions behind grid barriers, with contact
Fx = meshes.boundary_facets(T)
Pptr, Padj, _ = pbat.sim.xpbd.partition_mesh_constraints(X, T)
dx = (pbat.sim.xpbd.Data().with_volume_mesh(X, T).with_surface_mesh(np.unique(Fx), Fx)
      .with_bodies(np.zeros(X.shape[1], np.int64)).with_mass_inverse(np.full(X.shape[1], 0.1))
      .with_dirichlet_constrained_vertices(dbc).with_partitions(Pptr, Padj).construct())
xp = pbat.gpu.xpbd.Integrator(dx)
xp.step(0.01, 3, 2)
assert np.isfinite(xp.x).all()
# contact
Xb, Tb = meshes.tet_grid(2, 2, 1, 0.5)
Xt, Tt = meshes.tet_grid(1, 1, 1, 0.5, origin=(0.2, 0.3, 0.53))
Xc = np.concatenate([Xb, Xt], axis=1)
Tc = np.concatenate([Tb, Tt + Xb.shape[1]], axis=1)
B = np.concatenate([np.zeros(Xb.shape[1], np.int64), np.ones(Xt.shape[1], np.int64)])
F = meshes.boundary_facets(Tc)
V = np.unique(F)
v = np.zeros_like(Xc)
v[2, Xb.shape[1]:] = -1.0
d = (pbat.sim.vbd.Data().with_volume_mesh(Xc, Tc).with_surface_mesh(V, F).with_bodies(B).with_velocity(v)
     .with_dirichlet_vertices(np.flatnonzero(Xc[2] == 0)).construct())
vbd = pbat.gpu.vbd.Integrator(d)
for _ in range(4):
    vbd.step(0.01, 3, 1)
assert np.isfinite(vbd.x).all()
aabbs = pbat.gpu.geometry.Aabb()
aabbs.construct(Xc, F)
bvh = pbat.gpu.geometry.Bvh(F.shape[1], 64 * F.shape[1])
bvh.build(aabbs, Xc.min(axis=1), Xc.max(axis=1))
bvh.detect_overlaps(aabbs)
bvh.point_triangle_nearest_neighbours(aabbs, Xc[:, :5], Xc, F)
# greedy colouring on the device (FirstAvailable)
assert np.array_equal(pbat.graph.mesh_greedy_color(T, X.shape[1], 2, 1, device=0), pbat.graph.mesh_greedy_color(T, X.shape[1], 2, 1))
print("sanitize_case: done")
