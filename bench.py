#!/usr/bin/env python
"""Benchmark of the VBD hot path (BASELINE.json metric: VBD vertex-iterations/sec).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one ``Step(dt, iterations, substeps)`` of the integrator over the workload:
BASELINE.json configs[1] -- synthetic 58^3-cube tet grid (975,560 tets / 205,379 vertices),
Stable Neo-Hookean, 30 iterations/step, Chebyshev rho = 0.9, z = 0 face fixed.

  value      vertex-iterations/s, state resident in HBM, device-timed (CUDA events), max over ranks
  e2e        same metric through the public Python/C-ABI call with HOST buffers: every step
             uploads positions from pinned host memory, steps, and reads positions back
  roofline   algorithmic bytes per launch (B = k*68 + n*12 + 36 + 48) / kernel time
  cpu_baseline  the reference's CPU arithmetic (oracle/_ref, OpenMP over each colour) on this host

N > 1: one mesh domain-decomposed over the GPUs: a beam of N x 58^3 cubes, one 58^3 slab
(the N = 1 workload) per rank, so the per-GPU work is fixed ("weak" scaling).  After every colour the
owners push the new positions of the slab interfaces into their neighbours' ghost slots with peer-to-peer
stores over NVLink inside the step kernel; sweeps are barrier-free (values carry their write number),
only the barriers around the pre-step span the GPUs; NCCL (torch.distributed) only does the set-up exchange and the
timing reduction.  --replicas runs N independent copies instead.
--impl reference times the CPU reference on rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = 58
ITERS = 30
RHO = 0.9
DT = 0.01


def workload(seed=0, slabs=1):
    from physicsbasedanimationtoolkit_b200 import meshes

    X, T = meshes.tet_grid(GRID * slabs, GRID, GRID, 1.0 / GRID)
    dbc = np.flatnonzero(X[2] == 0)
    rng = np.random.default_rng(seed)
    x0 = X + 0.05 / GRID * rng.uniform(-1, 1, X.shape)
    x0[:, dbc] = X[:, dbc]
    return X, T, dbc, x0


def algorithmic_bytes_per_vertex_iteration(T, nV, active, chebyshev=True):
    """B = kbar*68 + nbar*12 + 36 (+48 Chebyshev), kbar = incident tets and nbar =
    1-ring size incl. self, both averaged over the swept vertices."""
    deg = np.bincount(T.reshape(-1), minlength=nV)
    pairs = np.concatenate([np.stack([T[a], T[b]]) for a in range(4) for b in range(4)], axis=1)
    key = np.unique(pairs[0].astype(np.int64) * nV + pairs[1])
    ring = np.bincount((key // nV).astype(np.int64), minlength=nV)  # includes self
    kbar = deg[active].mean()
    nbar = ring[active].mean()
    return float(kbar * 68 + nbar * 12 + 36 + (48 if chebyshev else 0)), float(kbar), float(nbar)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                clk, mx = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            smax = max(smax, mx)
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(clk)
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        if not sm:
            sm = [float(r[1].split(",")[0]) for r in self.rows[-3:] if r[1].split(",")[0].strip().replace(".", "").isdigit()] or [0.0]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference(X, T, dbc, x0, steps, warmup=0, threads=None):
    """Times the reference's CPU VBD arithmetic on this host.  Returns (value, info)."""
    import oracle

    kind = "reference" if oracle.have_ref() else "port"
    o = oracle.Oracle(X, T, dbc=dbc, accel=oracle.ACCEL_CHEBYSHEV, rho=RHO, kind=kind)
    if threads:
        o.set_num_threads(threads)
    cores = o.num_threads
    o.x = x0
    n_active = o.get("Padj").size
    for _ in range(warmup):
        o.step(DT, ITERS, 1)
    t = time.perf_counter()
    for _ in range(steps):
        o.step(DT, ITERS, 1)
    el = time.perf_counter() - t
    value = n_active * ITERS * steps / el
    info = {"value": value, "unit": "vertex-iterations/s", "cores": cores, "kind": kind,
            "sample": f"{steps} full step(s) of the same workload ({ITERS} iterations, Chebyshev) in {el:.1f} s; "
                      f"double precision, OpenMP over each colour, {cores} threads"}
    return value, el / steps, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=10, help="steps of the CPU-baseline sample (about 1.1 s each on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tile-iters", type=int, default=0)
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent replicas instead of domain decomposition")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = args.steps, max(args.warmup, 0)

    config = {"workload": f"configs[1]: synthetic {GRID}^3-cube tet grid, Stable Neo-Hookean, {ITERS} iterations/step, "
                          f"Chebyshev rho={RHO}, dt={DT}, z=0 face Dirichlet",
              "tets": None, "vertices": None, "iterations": ITERS, "substeps": 1,
              "parallelism": "single GPU" if args.gpus == 1 else (
                  f"{args.gpus} independent scene replicas (no collective)" if args.replicas else
                  f"domain decomposition: {args.gpus} x-slabs of {GRID}^3 cubes of one {GRID * args.gpus}x{GRID}x{GRID} beam, "
                  "per-colour halo push over NVLink (peer-to-peer stores inside the step kernel); halo and local values synchronise by write number, inter-GPU barriers only around the pre-step"),
              "l2": "working set per sweep (incidence-record stream) exceeds the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        # the arm's own workload: one 58^3 slab per GPU of the domain-decomposed beam (N slabs), or one replica
        slabs = 1 if args.replicas else max(args.gpus, 1)
        X, T, dbc, x0 = workload(slabs=slabs)
        config["tets"], config["vertices"] = int(T.shape[1]), int(X.shape[1])
        # bounded sample: about 25 s of CPU work (a step of one slab takes about 1.1 s on 16 cores)
        k = max(1, min(max(steps, 1), 24 // slabs))
        value, spstep, info = cpu_reference(X, T, dbc, x0, steps=k, warmup=min(warmup, 1))
        info["sample"] = f"bounded: {k} timed step(s) instead of {steps}; " + info["sample"]
        line = {"impl": "reference", "metric": "VBD vertex-iterations/sec", "value": value, "unit": "vertex-iterations/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": spstep * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "cpu_baseline": info,
                "e2e": {"value": value, "unit": "vertex-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch

    import physicsbasedanimationtoolkit_b200 as pbat

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    decomposed = world > 1 and not args.replicas
    if decomposed:
        from physicsbasedanimationtoolkit_b200.dist import DomainDecomposedIntegrator

        Xg, Tg, dbc_g, x0g = workload(seed=0, slabs=world)
        config["tets"], config["vertices"] = int(Tg.shape[1]), int(Xg.shape[1])
        os.environ.setdefault("VBDX_DIST_TIMEOUT_S", "10")  # the ranks start together here: a peer that is 10 s late is gone
        dd = DomainDecomposedIntegrator(Xg, Tg, dbc=dbc_g, rho_chebyshev=RHO, axis=0, tile_iters=args.tile_iters)
        # The sweeps are barrier-free across the GPUs; every dependency wait carries a time-out.  Probe
        # two steps; should any rank see a time-out, ALL ranks rebuild with colour barriers instead of reporting nothing.
        ok = 1
        try:
            dd.vbd.x = np.ascontiguousarray(x0g[:, dd.local.l2g], dtype=np.float32)
            dist.barrier()
            for _ in range(2):
                dd.vbd.step(DT, ITERS, 1)
        except RuntimeError as e:
            print(f"bench.py[{rank}]: barrier-free sweep failed ({e})", file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            os.environ["VBDX_DATAFLOW"] = "0"
            del dd
            dist.barrier()
            dd = DomainDecomposedIntegrator(Xg, Tg, dbc=dbc_g, rho_chebyshev=RHO, axis=0, tile_iters=args.tile_iters)
        vbd, lp = dd.vbd, dd.local
        X, T, dbc = lp.X, lp.T, np.concatenate([lp.dbc, lp.ghost_local])
        x0 = x0g[:, lp.l2g]
        nV, nT = X.shape[1], T.shape[1]
        info = vbd.info
        n_active = info["nActiveVertices"]
        n_active_job = int(Xg.shape[1] - dbc_g.size)
        del Xg, Tg, x0g
    else:
        X, T, dbc, x0 = workload(seed=rank)
        nV, nT = X.shape[1], T.shape[1]
        config["tets"], config["vertices"] = int(nT), int(nV)
        data = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc)
                .with_chebyshev_acceleration(RHO).construct())
        vbd = pbat.gpu.vbd.Integrator(data, device=local_rank, tile_iters=args.tile_iters)
        try:
            # The default sweep is barrier-free.  Its dependency waits carry a time-out that turns a
            # bug into an exception; should that ever fire, measure the barrier sweep instead of reporting nothing.
            vbd.x = np.ascontiguousarray(x0, dtype=np.float32)
            vbd.step(DT, ITERS, 1)
        except RuntimeError as e:
            print(f"bench.py: barrier-free sweep failed ({e}); falling back to the barrier sweep", file=sys.stderr)
            os.environ["VBDX_DATAFLOW"] = "0"
            vbd = pbat.gpu.vbd.Integrator(data, device=local_rank, tile_iters=args.tile_iters)
        config["sweep"] = "barrier" if os.environ.get("VBDX_DATAFLOW") == "0" else "barrier-free (dataflow-synchronised colours)"
        info = vbd.info
        n_active = info["nActiveVertices"]
        n_active_job = n_active * world
    config.setdefault("sweep", "barrier" if os.environ.get("VBDX_DATAFLOW") == "0" else "barrier-free (dataflow-synchronised colours)")
    active = np.ones(nV, bool)
    active[dbc] = False
    B, kbar, nbar = algorithmic_bytes_per_vertex_iteration(T, nV, active)

    x_start = np.ascontiguousarray(x0, dtype=np.float32)
    v_zero = np.zeros((3, nV), np.float32)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident leg: K steps back to back, per-step kernel time from CUDA events on the
    # launching stream (recorded inside the library around the cooperative launch)
    vbd.x = x_start
    vbd.v = v_zero
    barrier()  # domain decomposition: the ranks' kernels wait for each other, start them together
    for _ in range(warmup):
        vbd.step(DT, ITERS, 1)
    launches0 = vbd.info["kernelLaunches"]
    sampler = ClockSampler(local_rank)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.Stream()  # events must sit on the stream the kernels are launched on
    vbd.use_stream(stream.cuda_stream)
    if decomposed:
        vbd.dist_stats()  # reset the halo diagnostics
    t0 = time.time()
    ev0.record(stream)
    for _ in range(steps):
        vbd.step_async(DT, ITERS, 1)
    ev1.record(stream)
    vbd.synchronize()
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    total_ms = ev0.elapsed_time(ev1)
    halo = vbd.dist_stats() if decomposed else None
    launches = vbd.info["kernelLaunches"] - launches0
    # per-launch kernel duration (events around a single launch)
    kms = []
    for _ in range(min(steps, 5)):
        vbd.step(DT, ITERS, 1)
        kms.append(vbd.info["lastStepMs"])
    kernel_ms = float(np.mean(kms))
    assert np.isfinite(vbd.x).all(), "non-finite positions after the timed steps"

    # ---- end-to-end leg through the public API with HOST buffers (page-locked, pbat.host.pinned_empty):
    # every step uploads the input positions, steps, and reads the result back
    vbd.x = x_start
    vbd.v = v_zero
    xin = pbat.host.pinned_empty((3, nV), np.float32)
    xout = pbat.host.pinned_empty((3, nV), np.float32)
    xin[...] = x_start
    for _ in range(min(warmup, 2)):
        vbd.x = xin
        vbd.step(DT, ITERS, 1)
        vbd.positions(out=xout)
        xin, xout = xout, xin
    barrier()
    te = time.perf_counter()
    for _ in range(steps):
        vbd.x = xin                 # H2D of this step's input positions
        vbd.step(DT, ITERS, 1)
        vbd.positions(out=xout)     # D2H of the step's result
        xin, xout = xout, xin       # the result is the next step's input
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - te

    times = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(times[0]), float(times[1])
    work = n_active_job * ITERS * steps
    value = work / (total_ms * 1e-3)
    e2e_value = work / (e2e_ms * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        bytes_per_launch = n_active * ITERS * B
        achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        line = {
            "metric": "VBD vertex-iterations/sec", "value": value, "unit": "vertex-iterations/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "steps_per_s": steps * world / (total_ms * 1e-3),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "vertex-iterations/s", "h2d_bytes_per_step": int(nV * 12),
                    "d2h_bytes_per_step": int(nV * 12), "ms_per_step": e2e_ms / steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "vbdx::StepKernelPipe<true,false,false,%s> (one persistent cooperative launch per step)" % ("false" if os.environ.get("VBDX_DATAFLOW") == "0" else "true"),
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": bytes_per_launch,
                         "bytes_per_vertex_iteration": B, "kbar": kbar, "nbar": nbar,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6.65 TB/s",
                         "streamed_record_bytes_per_launch": int(info["nRecordSlots"] * 32 * ITERS),
                         # what the DRAM counters saw (ncu, profiles/traffic.json) over the live kernel time: the closed-form
                         # records stream less than half of the algorithmic bytes, which is how frac can exceed 1
                         "traffic_frac": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None},
        }
        if halo is not None:
            line["halo"] = dict(halo, note="rank 0, over the timed steps: ghost values that had to be polled / ns polling (summed over lanes) / barriers of CTA 0 that waited for a neighbour's epoch / ns")
        if world == 1 and not args.no_cpu_baseline:
            _, _, cpu = cpu_reference(X, T, dbc, x0, steps=args.cpu_steps)
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
