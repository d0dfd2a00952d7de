"""Barrier-free (dataflow) sweeps against the barrier sweeps: bit-identical results, ms/step.
  python tools/dataflow_check.py [grid]"""
import os, subprocess, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 2 and sys.argv[2] == "child":
    import numpy as np
    import physicsbasedanimationtoolkit_b200 as pbat
    from physicsbasedanimationtoolkit_b200 import meshes
    n = int(sys.argv[1])
    X, T = meshes.tet_grid(n, n, n, 1.0 / n)
    dbc = np.flatnonzero(X[2] == 0)
    x0 = X + 0.05 / n * np.random.default_rng(0).uniform(-1, 1, X.shape)
    x0[:, dbc] = X[:, dbc]
    out = {}
    for cheb in (0.9, None):
        d = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc)
        if cheb:
            d = d.with_chebyshev_acceleration(cheb)
        d = d.construct()
        vbd = pbat.gpu.vbd.Integrator(d, kernel_variant=3)
        vbd.x = x0.astype(np.float32)
        ms = []
        for _ in range(12):
            vbd.step(0.01, 30, 2 if cheb is None else 1)
            ms.append(vbd.info["lastStepMs"])
        x = vbd.x
        out["cheb" if cheb else "base"] = {"ms": float(np.median(ms[2:])), "hash": __import__("hashlib").sha1(x.tobytes()).hexdigest(), "finite": bool(np.isfinite(x).all())}
    print(json.dumps(out))
else:
    grid = sys.argv[1] if len(sys.argv) > 1 else "58"
    res = {}
    for mode in ("0", "1"):
        env = dict(os.environ, VBDX_DATAFLOW=mode, VBDX_DATAFLOW_TIMEOUT_S="3")
        r = subprocess.run([sys.executable, __file__, grid, "child"], env=env, capture_output=True, text=True, timeout=50)
        print(f"dataflow={mode}:", r.stdout.strip()[-400:], r.stderr.strip()[-600:])
        try:
            res[mode] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:  # noqa: BLE001
            res[mode] = None
    if res["0"] and res["1"]:
        for k in res["0"]:
            print(k, "bit-identical:", res["0"][k]["hash"] == res["1"][k]["hash"], "ms barrier", res["0"][k]["ms"], "dataflow", res["1"][k]["ms"])
