"""CPU oracle for the VBD hot path -- TEST INFRASTRUCTURE ONLY.

ctypes front-end over ``liboracle_port.so`` (restated arithmetic) and, when present,
``_ref/liboracle_ref.so`` (per-vertex arithmetic compiled from the reference's own headers).
See ``vbd_oracle.cpp`` for the reference file:line each routine follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(_HERE, "liboracle_port.so")
REF_LIB = os.path.join(_HERE, "_ref", "liboracle_ref.so")

# enums shared with the product (sim/vbd/Enums.h:9-28, graph/Enums.h)
POSITION, INERTIA, KINETIC_ENERGY_MINIMUM, ADAPTIVE_VBD, ADAPTIVE_PBAT = range(5)
ACCEL_NONE, ACCEL_CHEBYSHEV, ACCEL_ANDERSON, ACCEL_NESTEROV, ACCEL_BROYDEN, ACCEL_TRUST_REGION = 0, 1, 2, 3, 4, 5   # = vbdx_acceleration_strategy
ORDER_NATURAL, ORDER_SMALLEST_DEGREE, ORDER_LARGEST_DEGREE = range(3)
SELECT_LEAST_USED, SELECT_FIRST_AVAILABLE = range(2)
MATERIAL_STABLE_NEO_HOOKEAN, MATERIAL_STVK = range(2)  # = vbdx_material


def build(ref: bool = True) -> None:
    """Compile the oracle(s). ``_ref`` is only (re)built where /root/reference exists."""
    targets = ["port"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", _HERE, "-s"] + targets, check=True)


class _Desc(C.Structure):
    _fields_ = [
        ("nV", C.c_int64), ("nT", C.c_int64),
        ("X", C.c_void_p), ("E", C.c_void_p), ("v", C.c_void_p), ("aext", C.c_void_p),
        ("rhoe", C.c_void_p), ("lame", C.c_void_p), ("dbc", C.c_void_p), ("nDbc", C.c_int64),
        ("colors", C.c_void_p),
        ("ordering", C.c_int), ("selection", C.c_int),
        ("strategy", C.c_int), ("accel", C.c_int), ("omegaMode", C.c_int),
        ("kD", C.c_double), ("detHZero", C.c_double), ("rho", C.c_double),
        ("B", C.c_void_p), ("V", C.c_void_p), ("nCV", C.c_int64), ("F", C.c_void_p), ("nF", C.c_int64),
        ("muC", C.c_double), ("muF", C.c_double), ("epsv", C.c_double), ("activeSetUpdateFrequency", C.c_int64),
    ]


_libs: dict[str, C.CDLL] = {}


def _load(kind: str) -> C.CDLL:
    if kind in _libs:
        return _libs[kind]
    path = PORT_LIB if kind == "port" else REF_LIB
    if not os.path.exists(path):
        if kind == "port":
            build(ref=False)
        else:
            raise FileNotFoundError(f"{path} not built (needs /root/reference; run `make -C oracle ref`)")
    lib = C.CDLL(path)
    lib.vbdo_kind.restype = C.c_char_p
    lib.vbdo_last_error.restype = C.c_char_p
    lib.vbdo_create.restype = C.c_void_p
    lib.vbdo_create.argtypes = [C.POINTER(_Desc)]
    lib.vbdo_destroy.argtypes = [C.c_void_p]
    lib.vbdo_step.argtypes = [C.c_void_p, C.c_double, C.c_int64, C.c_int64]
    lib.vbdo_sweeps.argtypes = [C.c_void_p, C.c_double, C.c_int64]
    lib.vbdo_size.restype = C.c_int64
    lib.vbdo_size.argtypes = [C.c_void_p, C.c_char_p]
    lib.vbdo_get_f64.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    lib.vbdo_get_i64.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    lib.vbdo_set_f64.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    lib.vbdo_set_params.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
    lib.vbdo_set_acceleration.restype = C.c_int
    lib.vbdo_set_acceleration.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int64, C.c_int64]
    lib.vbdo_xpbd_setup.restype = C.c_int
    lib.vbdo_xpbd_setup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vbdo_xpbd_step.argtypes = [C.c_void_p, C.c_double, C.c_int64, C.c_int64]
    lib.vbdo_set_trust_region.restype = C.c_int
    lib.vbdo_set_trust_region.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
    lib.vbdo_objective.restype = C.c_double
    lib.vbdo_objective.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
    lib.vbdo_objective_gradient.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    lib.vbdo_snh_eval.restype = C.c_double
    lib.vbdo_snh_eval.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.vbdo_stvk_eval.restype = C.c_double
    lib.vbdo_stvk_eval.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.vbdo_set_material.argtypes = [C.c_void_p, C.c_int]
    lib.vbdo_set_line_search_guard.argtypes = [C.c_void_p, C.c_int]
    lib.vbdo_num_threads.restype = C.c_int
    lib.vbdo_set_num_threads.argtypes = [C.c_int]
    _libs[kind] = lib
    return lib


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def default_kind() -> str:
    """The checker the tests use when they do not name one: the build over the reference's own headers where it exists
    (``oracle/_ref`` travels to the GPU box with the repository snapshot), else the port.  ``VBD_ORACLE_KIND`` overrides."""
    k = os.environ.get("VBD_ORACLE_KIND", "")
    if k in ("port", "reference"):
        return k
    return "reference" if have_ref() else "port"


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_F64 = ("x", "v", "aext", "xt", "xtilde", "vt", "X", "m", "GP", "wg", "lame")
_I64 = ("GVGp", "GVGe", "GVGilocal", "colors", "Pptr", "Padj", "nn", "fc", "active")


class Oracle:
    """Double-precision CPU VBD integrator with the reference's semantics.

    ``X`` is 3 x nV, ``E`` is 4 x nT (the reference's column-per-vertex / column-per-tet
    convention, sim/vbd/Data.h:167-170).  ``kind`` = "port" | "reference" (default: ``default_kind()``).
    """

    def __init__(self, X, E, *, v=None, aext=None, rhoe=None, mue=None, lambdae=None, dbc=None,
                 colors=None, ordering=ORDER_LARGEST_DEGREE, selection=SELECT_LEAST_USED,
                 strategy=ADAPTIVE_PBAT, accel=ACCEL_NONE, rho=1.0, omega_mode=0, kD=0.0,
                 detH_zero=1e-7, kind=None, B=None, V=None, F=None, muC=1e6, muF=0.3, epsv=1e-3,
                 active_set_update_frequency=1, material=MATERIAL_STABLE_NEO_HOOKEAN):
        kind = kind or default_kind()
        self.lib = _load("port" if kind == "port" else "ref")
        X = np.asarray(X, dtype=np.float64)
        E = np.asarray(E, dtype=np.int64)
        assert X.shape[0] == 3 and E.shape[0] == 4
        self.nV, self.nT = X.shape[1], E.shape[1]
        keep = []

        def colmajor(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(np.asarray(a, dtype=dt).T)  # (n, rows): column-major of rows x n
            keep.append(a)
            return a

        Xc, Ec = colmajor(X, np.float64), colmajor(E, np.int64)
        vc, ac = colmajor(v, np.float64), colmajor(aext, np.float64)
        lame = None
        if mue is not None:
            lame = np.ascontiguousarray(np.stack([np.asarray(mue, np.float64),
                                                  np.asarray(lambdae, np.float64)], axis=1))
        rhoe = None if rhoe is None else np.ascontiguousarray(rhoe, dtype=np.float64)
        dbc = None if dbc is None else np.ascontiguousarray(dbc, dtype=np.int64)
        colors = None if colors is None else np.ascontiguousarray(colors, dtype=np.int64)
        B = None if B is None else np.ascontiguousarray(B, dtype=np.int64)
        V = None if V is None else np.ascontiguousarray(V, dtype=np.int64)
        Fc = None if F is None else np.ascontiguousarray(np.asarray(F, dtype=np.int64).T)
        d = _Desc(self.nV, self.nT, _ptr(Xc), _ptr(Ec), _ptr(vc), _ptr(ac), _ptr(rhoe), _ptr(lame),
                  _ptr(dbc), 0 if dbc is None else dbc.size, _ptr(colors), ordering, selection,
                  strategy, accel, omega_mode, kD, detH_zero, rho,
                  _ptr(B), _ptr(V), 0 if V is None else V.size, _ptr(Fc), 0 if Fc is None else Fc.shape[0],
                  muC, muF, epsv, active_set_update_frequency)
        self.nCV = 0 if V is None else V.size
        self.h = self.lib.vbdo_create(C.byref(d))
        if not self.h:
            raise ValueError(self.lib.vbdo_last_error().decode())
        self.kind = self.lib.vbdo_kind().decode()
        self.lib.vbdo_set_material(self.h, int(material))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.vbdo_destroy(self.h)
            self.h = None

    def step(self, dt, iterations, substeps=1):
        self.lib.vbdo_step(self.h, dt, iterations, substeps)

    def sweeps(self, dt, reps):
        self.lib.vbdo_sweeps(self.h, dt, reps)

    def get(self, name):
        n = self.lib.vbdo_size(self.h, name.encode())
        if name in _F64:
            out = np.empty(n, dtype=np.float64)
            rc = self.lib.vbdo_get_f64(self.h, name.encode(), _ptr(out))
        elif name in _I64:
            out = np.empty(n, dtype=np.int64)
            rc = self.lib.vbdo_get_i64(self.h, name.encode(), _ptr(out))
        else:
            raise KeyError(name)
        assert rc == 0
        if name in ("x", "v", "aext", "xt", "xtilde", "vt", "X"):
            return out.reshape(self.nV, 3).T.copy()
        if name == "GP":
            # 4 x 3nT, element e = columns 3e..3e+2 (sim/vbd/Data.h:191-192)
            return out.reshape(3 * self.nT, 4).T.copy()
        if name == "lame":
            return out.reshape(self.nT, 2).T.copy()
        if name == "nn":
            return out.reshape(-1, 8)
        if name == "fc":
            return out.reshape(-1, 8)
        return out

    def set(self, name, a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).T)
        assert self.lib.vbdo_set_f64(self.h, name.encode(), _ptr(a)) == 0

    def set_line_search_guard(self, on):
        """The product's guarded Newton step (include/vbdx.h vbdx_set_line_search_guard), restated; off = reference."""
        self.lib.vbdo_set_line_search_guard(self.h, 1 if on else 0)

    def set_params(self, strategy, kD, detH_zero):
        self.lib.vbdo_set_params(self.h, strategy, kD, detH_zero)

    x = property(lambda s: s.get("x"), lambda s, a: s.set("x", a))
    v = property(lambda s: s.get("v"), lambda s, a: s.set("v", a))

    def set_acceleration(self, accel, rho=1.0, L=1.0, start=3, window=5):
        """Data::With{Chebyshev,Anderson,Nesterov}Acceleration; ``accel`` = ACCEL_* (the reference validates in
        Data::Construct, sim/vbd/Data.cpp:268-296)."""
        if self.lib.vbdo_set_acceleration(self.h, int(accel), float(rho), float(L), int(start), int(window)):
            raise ValueError("invalid acceleration parameters")

    def xpbd_setup(self, Pptr, Padj, minv=None, muV=None, muS=0.3, muD=0.2, beta_snh=None, alpha_c=None, beta_c=None):
        """Turns this problem into an XPBD one (sim/xpbd/Data.cpp:102-160): constraint partitions ``Pptr`` / ``Padj`` over the
        tets, inverse masses (default 1e-3), collision penalties per collision vertex (default 1), friction coefficients,
        optional damping of the elastic constraints (2 per tet) and compliance / damping of the contacts."""
        def arr(a, dt=np.float64):
            return None if a is None else np.ascontiguousarray(a, dtype=dt)
        Pptr, Padj = arr(Pptr, np.int64), arr(Padj, np.int64)
        minv, muV, beta_snh, alpha_c, beta_c = arr(minv), arr(muV), arr(beta_snh), arr(alpha_c), arr(beta_c)
        if self.lib.vbdo_xpbd_setup(self.h, _ptr(minv), _ptr(muV), float(muS), float(muD), _ptr(Pptr), Pptr.size - 1, _ptr(Padj),
                                    _ptr(beta_snh), _ptr(alpha_c), _ptr(beta_c)):
            raise ValueError(self.lib.vbdo_last_error().decode())

    def xpbd_step(self, dt, iterations, substeps=1):
        """sim/xpbd/Integrator.cpp:31-139 with the GPU path's contact pipeline (gpu/impl/xpbd/Integrator.cu:88-186)."""
        self.lib.vbdo_xpbd_step(self.h, float(dt), int(iterations), int(substeps))

    def set_trust_region(self, eta=0.2, tau=2.0, curved=True):
        """Data::WithTrustRegionAcceleration (sim/vbd/Data.cpp); the solve follows gpu/impl/vbd/TrustRegionIntegrator.cu."""
        if self.lib.vbdo_set_trust_region(self.h, float(eta), float(tau), int(bool(curved))):
            raise ValueError("Expected eta >= 0 and tau > 1")

    def objective(self, xk, xtilde, dt):
        a = np.ascontiguousarray(np.asarray(xk, np.float64).T)
        b = np.ascontiguousarray(np.asarray(xtilde, np.float64).T)
        return self.lib.vbdo_objective(self.h, _ptr(a), _ptr(b), dt)

    def objective_gradient(self, xk, xtilde, dt):
        a = np.ascontiguousarray(np.asarray(xk, np.float64).T)
        b = np.ascontiguousarray(np.asarray(xtilde, np.float64).T)
        out = np.empty(3 * self.nV)
        self.lib.vbdo_objective_gradient(self.h, _ptr(a), _ptr(b), dt, _ptr(out))
        return out

    @property
    def num_threads(self):
        return self.lib.vbdo_num_threads()

    def set_num_threads(self, n):
        self.lib.vbdo_set_num_threads(n)


def snh_eval(F, mu, lam, kind="port"):
    lib = _load("port" if kind == "port" else "ref")
    f = np.ascontiguousarray(np.asarray(F, np.float64).T).reshape(-1)  # column-major
    return lib.vbdo_snh_eval(_ptr(f), mu, lam)


def stvk_eval(F, mu, lam, kind="port"):
    lib = _load("port" if kind == "port" else "ref")
    f = np.ascontiguousarray(np.asarray(F, np.float64).T).reshape(-1)  # column-major
    return lib.vbdo_stvk_eval(_ptr(f), mu, lam)
