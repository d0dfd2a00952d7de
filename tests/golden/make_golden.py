"""Generates tests/golden/vbd_golden.npz from the *reference-header* oracle build
(oracle/_ref/liboracle_ref.so: the reference's own sim/vbd/Kernels.h and
physics/StableNeoHookeanEnergy.h compiled where they lie under /root/reference).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The fixtures pin (a) the restated oracle port and (b) the CUDA path, on any machine.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from physicsbasedanimationtoolkit_b200 import meshes  # noqa: E402

CASES = {
    # name: (mesh, dict(oracle kwargs), dt, iterations, substeps, steps)
    "cube_base": (("cube",), {}, 1e-2, 10, 1, 1),
    "cube_cheb": (("cube",), dict(accel=1, rho=0.9), 1e-2, 10, 1, 1),
    "beam_small_base": (("grid", 6, 3, 3, 0.1), {}, 1e-2, 10, 1, 10),
    "beam_small_cheb": (("grid", 6, 3, 3, 0.1), dict(accel=1, rho=0.9), 1e-2, 10, 1, 10),
    "beam_small_cheb_textbook": (("grid", 6, 3, 3, 0.1), dict(accel=1, rho=0.9, omega_mode=1), 1e-2, 10, 1, 10),
    "beam_small_substeps_damped": (("grid", 6, 3, 3, 0.1), dict(kD=1e-3), 2e-2, 8, 3, 5),
    "beam_small_position": (("grid", 6, 3, 3, 0.1), dict(strategy=0), 1e-2, 10, 1, 10),
    "beam_small_inertia": (("grid", 6, 3, 3, 0.1), dict(strategy=1), 1e-2, 10, 1, 10),
    "beam_small_kinetic": (("grid", 6, 3, 3, 0.1), dict(strategy=2), 1e-2, 10, 1, 10),
    "beam_small_adaptive_vbd": (("grid", 6, 3, 3, 0.1), dict(strategy=3), 1e-2, 10, 1, 10),
    # St. Venant-Kirchhoff in place of the Stable Neo-Hookean energy (reference header physics/SaintVenantKirchhoffEnergy.h)
    "beam_small_stvk": (("grid", 6, 3, 3, 0.1), dict(material=1), 1e-2, 10, 1, 10),
    "beam_small_stvk_cheb_damped": (("grid", 6, 3, 3, 0.1), dict(material=1, accel=1, rho=0.9, kD=1e-3), 1e-2, 10, 2, 10),
    "config1_base": (("grid", 25, 9, 9, 0.04), {}, 1e-2, 20, 1, 100),
    "config1_cheb": (("grid", 25, 9, 9, 0.04), dict(accel=1, rho=0.9), 1e-2, 20, 1, 100),
}


def mesh_of(spec):
    if spec[0] == "cube":
        return meshes.CUBE_P, meshes.CUBE_T, None
    X, T = meshes.tet_grid(*spec[1:4], spec[4])
    return X, T, np.flatnonzero(X[0] == 0)


def run(name, kind):
    spec, kw, dt, iters, sub, steps = CASES[name]
    X, T, dbc = mesh_of(spec)
    o = oracle.Oracle(X, T, dbc=dbc, kind=kind, **kw)
    for _ in range(steps):
        o.step(dt, iters, sub)
    return o.x, o.v, o.get("colors")


if __name__ == "__main__":
    assert oracle.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for name in CASES:
        x, v, c = run(name, "reference")
        out[name + "/x"] = x
        out[name + "/v"] = v
        out[name + "/colors"] = c.astype(np.int16)
        print(name, x.shape, float(np.abs(x).max()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vbd_golden.npz"), **out)
