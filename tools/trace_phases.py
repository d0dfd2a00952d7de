"""Diagnostics: where does the time of a colour phase go (direct kernel)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import meshes
n = 58
X, T = meshes.tet_grid(n, n, n, 1.0 / n)
dbc = np.flatnonzero(X[2] == 0)
data = pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_chebyshev_acceleration(0.9).construct()
for ti, variant, cw in ((8, 3, 0),):
    vbd = pbat.gpu.vbd.Integrator(data, tile_iters=ti, kernel_variant=variant, consumer_warps=cw)
    for _ in range(3):
        vbd.step(0.01, 30, 1)
    tr = vbd.trace_phases(10, 0.01, 30, 1).astype(np.int64)
    t0 = tr[..., 0].min()
    tr = tr - t0
    print(f"variant={variant} cw={cw} tile_iters={ti} step {vbd.info['lastStepMs']:.3f} ms; iteration span {(tr[...,3].max())/1e3:.1f} us")
    for c in range(tr.shape[0]):
        s, w0, cta, rel = (tr[c, :, i] for i in range(4))
        has = tr[c, :, 4] > 0
        td_, st_, ac_, so_ = (np.median((tr[c, has, i] - tr[c, has, 0])) / 1e3 for i in (4, 5, 6, 7))
        if tr.shape[2] > 8 and (tr[c, :, 8] > 0).any():
            h8 = tr[c, :, 8] > 0
            a, b, cc, d = (np.median(tr[c, h8, i] - tr[c, h8, 2]) / 1e3 for i in (8, 9, 10, 11))
            f_ = np.median(tr[c, h8, 1] - tr[c, h8, 2]) / 1e3
            print(f"      warp0 shadow work (us since CTA arrival): fence+red {f_:.2f} waited {a:.2f} records {b:.2f} gather {cc:.2f} ids {d:.2f}")
        print(f"      warp0 first tile (median over CTAs, us since phase start): desc {td_:.2f} staged {st_:.2f} accumulated {ac_:.2f} solved {so_:.2f}")
        print(f" colour {c}: start {s.min()/1e3:7.2f}..{s.max()/1e3:7.2f}  cta work med {np.median(cta-s)/1e3:5.2f} max {(cta-s).max()/1e3:5.2f} us | "
              f"last cta done {cta.max()/1e3:7.2f} | release first {rel.min()/1e3:7.2f} last {rel.max()/1e3:7.2f} | release-after-last-arrival {(rel.min()-cta.max())/1e3:5.2f}..{(rel.max()-cta.max())/1e3:5.2f}")
