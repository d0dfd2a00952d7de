"""Scratch experiments on the config-2 mesh: pipeline flags (VBDX_PIPE_FLAGS), warps per CTA, tile sizes; per-tile anatomy of
the barrier-free sweep from %globaltimer stamps.  python tools/exp_flow.py [grid] ['[(flags, consumer_warps, tile_iters), ...]']"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 58
# (flow kernel?, VBDX_PIPE_FLAGS, consumer warps, tile_iters)
variants = eval(sys.argv[2]) if len(sys.argv) > 2 else [(1, 0, 0, 8), (0, 0, 0, 8), (0, 3, 0, 8), (1, 4, 0, 8), (1, 0, 15, 8), (1, 0, 0, 4), (1, 4, 0, 4), (1, 0, 0, 16)]
iters = 30
X, T = meshes.tet_grid(n, n, n, 1.0 / n)
dbc = np.flatnonzero(X[2] == 0)
data = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_chebyshev_acceleration(0.9).construct()
rng = np.random.default_rng(0)
xp = (X + 0.05 / n * rng.uniform(-1, 1, X.shape)).astype(np.float32)
xp[:, dbc] = X[:, dbc]
ref = {}
import hashlib
for flow, flags, cw, ti in variants:
    os.environ["VBDX_PIPE_FLAGS"] = str(flags)
    os.environ["VBDX_FLOW"] = str(flow)
    try:
        vbd = pbat.gpu.vbd.Integrator(data, tile_iters=ti, consumer_warps=cw)
    except Exception as e:
        print("skip", flow, flags, cw, ti, e, flush=True)
        continue
    info = vbd.info
    vbd.x = xp
    vbd.v = np.zeros_like(xp)
    for _ in range(3):
        vbd.step(0.01, iters, 1)
    ms = []
    for _ in range(12):
        vbd.step(0.01, iters, 1)
        ms.append(vbd.info["lastStepMs"])
    x = vbd.x
    key = ti
    same = None
    if key in ref:
        same = bool(np.array_equal(ref[key], x))
    else:
        ref[key] = x
    ms = np.array(ms)
    print(f"flow={flow} flags={flags} cw={cw} tile_iters={ti} sha={hashlib.sha1(x.tobytes()).hexdigest()[:10]}: grid={info['gridBlocks']}x{info['blockThreads']} tiles={info['nTiles']} step ms min/med={ms.min():.4f}/{np.median(ms):.4f} "
          f"-> {info['nActiveVertices']*iters/(np.median(ms)*1e-3)/1e9:.3f} Gvert-it/s  bit-identical to first of this tile size: {same} finite={np.isfinite(x).all()}", flush=True)
    if cw == 0 and ti == 8 and "--no-trace" not in sys.argv:
        tr = vbd.trace_phases(10, 0.01, iters, 1).astype(np.int64)     # [colour, CTA, 12]
        ok = (tr[..., 4] > 0) & (tr[..., 7] > 0)
        def med(a, b):
            d = (tr[..., a] - tr[..., b])[ok]
            return np.median(d) / 1e3, np.percentile(d, 90) / 1e3
        print(f"   anatomy of warp 0's tile per colour and CTA (us, median / p90) at {vbd.info['lastStepMs']:.3f} ms/step:")
        for name, a, b in (("start -> deps satisfied", 8, 4), ("deps -> records waited (loop start)", 5, 8), ("loop", 6, 5),
                           ("reduce + solve + store", 11, 6), ("late hook", 7, 11), ("whole tile", 7, 4)):
            m, p9 = med(a, b)
            print(f"     {name:40s} {m:6.2f} / {p9:6.2f}")
        nxt = tr[1:, :, 4] - tr[:-1, :, 7]
        okn = ok[1:] & ok[:-1]
        per = tr[1:, :, 4] - tr[:-1, :, 4]
        print(f"     {'tile end -> next tile start':40s} {np.median(nxt[okn])/1e3:6.2f} / {np.percentile(nxt[okn],90)/1e3:6.2f}")
        print(f"     {'period (start to next colour start)':40s} {np.median(per[okn])/1e3:6.2f} / {np.percentile(per[okn],90)/1e3:6.2f}")
        spread = [(tr[c, ok[c], 11].max() - tr[c, ok[c], 11].min()) / 1e3 for c in range(tr.shape[0])]
        print(f"     spread of the store time over CTAs per colour: {np.round(spread, 2)}")
    del vbd
