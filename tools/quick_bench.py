"""Scratch timing of the persistent step kernel on the config-2 mesh (not the judged bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 58
iters = 30
X, T = meshes.tet_grid(n, n, n, 1.0 / n)
rng = np.random.default_rng(0)
X0 = X.copy()
dbc = np.flatnonzero(X[2] == 0)
t = time.time()
data = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_chebyshev_acceleration(0.9).construct()
print(f"mesh {n}^3: nV={X.shape[1]} nT={T.shape[1]} construct {time.time()-t:.2f}s", flush=True)
configs = [(True, 1, 8, 0), (True, 2, 4, 0, 16)] + [(True, 3, ti, 0) for ti in (4, 8)] + [(False, 3, 8, 0)]
if len(sys.argv) > 2:
    configs = eval(sys.argv[2])
for cheb, variant, ti, rs, *rest in configs:
    cw = rest[0] if rest else 0
    data.accelerator = pbat.sim.vbd.AccelerationStrategy.Chebyshev if cheb else pbat.sim.vbd.AccelerationStrategy.Base
    if True:
        t = time.time()
        try:
            vbd = pbat.gpu.vbd.Integrator(data, tile_iters=ti, kernel_variant=variant, ring_slots=rs, consumer_warps=cw)
        except Exception as e:
            print('skip', cheb, variant, ti, rs, cw, e); continue
        tc = time.time() - t
        info = vbd.info
        xp = (X0 + 0.05 / n * rng.uniform(-1, 1, X.shape)).astype(np.float32)
        xp[:, dbc] = X0[:, dbc]
        vbd.x = xp
        for _ in range(3):
            vbd.step(0.01, iters, 1)
        ms = []
        for _ in range(10):
            vbd.step(0.01, iters, 1)
            ms.append(vbd.info["lastStepMs"])
        ms = np.array(ms)
        nact = info["nActiveVertices"]
        vips = nact * iters / (ms.min() * 1e-3)
        kbar = info["nIncidences"] / nact
        B = kbar * 68 + 12.7 * 12 + 36 + (48 if cheb else 0)
        print(f"cheb={cheb} variant={variant} ring={rs} cw={cw} tile_iters={ti}: create {tc:.2f}s grid={info['gridBlocks']}x{info['blockThreads']} tiles={info['nTiles']} "
              f"slots/inc={info['nRecordSlots']/info['nIncidences']:.3f} step ms min/med={ms.min():.3f}/{np.median(ms):.3f} "
              f"-> {vips/1e9:.3f} Gvert-it/s, {vips*B/1e9:.0f} GB/s algorithmic ({vips*B/6552e9:.2f} of 6552), "
              f"record stream {info['nRecordSlots']*32*iters/(ms.min()*1e-3)/1e9:.0f} GB/s", flush=True)
        assert np.isfinite(vbd.x).all()
        del vbd
