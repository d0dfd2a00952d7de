#!/usr/bin/env python
"""Benchmark of the VBD hot path (BASELINE.json metric: VBD vertex-iterations/sec).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one ``Step(dt, iterations, substeps)`` of the integrator over the workload:
BASELINE.json configs[1] -- synthetic 58^3-cube tet grid (975,560 tets / 205,379 vertices),
Stable Neo-Hookean, 30 iterations/step, Chebyshev rho = 0.9, z = 0 face fixed.

  value      vertex-iterations/s, state resident in HBM, device-timed (CUDA events), max over ranks
  e2e        same metric through the public Python/C-ABI call with HOST buffers: every step
             uploads positions from pinned host memory, steps, and reads positions back (enqueued with the
             asynchronous calls, one vbdx_synchronize per step: the host holds the result when the step returns)
  roofline   bytes the record format must move per launch / kernel time, with the SURVEY.md 8(d)
             figure beside it (frac_survey_formula) and the DRAM traffic ncu measured for this configuration
             (profiles/traffic.json, keyed by workload and GPU count)
  cpu_baseline  the reference's CPU arithmetic (oracle/_ref, OpenMP over each colour) on this host
  colouring_first_available   (N = 1) the same workload under the reference's other colour selection (4 colours instead of 7)

N > 1 (one process per GPU, torch.distributed/NCCL for set-up and timing only):
  main line  one mesh domain-decomposed over the GPUs: a beam of N x 58^3 cubes, one 58^3 slab (the N = 1
             workload) per rank -- per-GPU work fixed, "weak".  After every colour the owners push the new positions of the
             slab interfaces into their neighbours' ghost slots with peer-to-peer stores over NVLink inside the step kernel;
             sweeps are barrier-free (values carry their write number).
  parity_vs_single_gpu   the same N-slab beam stepped on ONE GPU (rank 0) from the same state: max |difference| (must be 0.0)
  strong     BASELINE.json configs[3]: ONE 117^3 block (8.0 M tets) domain-decomposed over the N GPUs, with the single-GPU
             time of the same block measured in the same run (rank 0) and the same bitwise comparison
  batch      BASELINE.json configs[4]: independent 10^3-cube scenes, 512 per GPU (4096 on 8 GPUs), no communication
--replicas runs N independent copies of configs[1] instead of the domain decomposition.
--impl reference times the CPU reference on rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = 58
STRONG_GRID = 117
ITERS = 30
RHO = 0.9
DT = 0.01
BATCH_SCENES_PER_GPU = 512
BATCH_ITERS = 20


def workload(seed=0, slabs=1, grid=GRID):
    from physicsbasedanimationtoolkit_b200 import meshes

    X, T = meshes.tet_grid(grid * slabs, grid, grid, 1.0 / grid)
    dbc = np.flatnonzero(X[2] == 0)
    rng = np.random.default_rng(seed)
    x0 = X + 0.05 / grid * rng.uniform(-1, 1, X.shape)
    x0[:, dbc] = X[:, dbc]
    return X, T, dbc, x0


def survey_bytes_per_vertex_iteration(kbar, nbar, chebyshev=True):
    """B = kbar*68 + nbar*12 + 36 (+48 Chebyshev), kbar = incident tets and nbar = 1-ring size incl.
    self, both averaged over the swept vertices."""
    return float(kbar * 68 + nbar * 12 + 36 + (48 if chebyshev else 0))


def format_bytes_per_vertex_iteration(info, chebyshev=True):
    """What THIS kernel's data format must move per vertex solve, from the handle's own counts:

      streamed from HBM every sweep (no reuse inside a sweep; 4 incidence records share no bytes):
        incidence records   32 B x incident tets                    (closed-form record: 3 packed ring indices + 6 scalars)
        ring entries         4 B x staged vertices of the tile      (pre-decoded gather index)
        tile descriptor     16 B x tiles
      per-vertex state, read or written once per solve (L2-resident at this size, counted all the same):
        xtilde+mass 16 B read, own position 16 B read, position write 16 B; Chebyshev: + P write 16, history read 16 + write 16
      served by the L2 (reported separately, not part of `achieved`): 16 B x staged vertices (the 1-ring gather)
    """
    n = float(info["nActiveVertices"])
    stream = (32.0 * info["nIncidences"] + 4.0 * info["nRingEntries"] + 16.0 * info["nTiles"]) / n
    state = 48.0 + (48.0 if chebyshev else 0.0)
    gather = 16.0 * info["nRingEntries"] / n
    return {"stream": stream, "state": state, "l2_gather": gather, "total": stream + state}


def mesh_stats(T, nV, active):
    deg = np.bincount(T.reshape(-1), minlength=nV)
    pairs = np.concatenate([np.stack([T[a], T[b]]) for a in range(4) for b in range(4)], axis=1)
    key = np.unique(pairs[0].astype(np.int64) * nV + pairs[1])
    ring = np.bincount((key // nV).astype(np.int64), minlength=nV)  # includes self
    return float(deg[active].mean()), float(ring[active].mean())


class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled through NVML from a thread every few milliseconds (the timed
    region lasts tens of milliseconds: nvidia-smi's own polling loop would not see it); falls back to nvidia-smi."""

    def __init__(self, index, period_s=0.002):
        self.rows = []
        self.stop_flag = False
        self.smax = 0.0
        self.thread = None
        self.index = index
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.period = period_s
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001 -- NVML missing: one nvidia-smi sample at stop()
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                clk = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.time(), clk, reasons))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        if self.nv is not None and self.rows:
            inside = [r for r in self.rows if t0 - 0.005 <= r[0] <= t1 + 0.005] or self.rows[-3:]
            reasons = set()
            for _, _, bits in inside:
                reasons.update(n for b, n in names.items() if bits & b)
            return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": self.smax, "reasons": sorted(reasons),
                    "samples": len(inside), "source": "NVML, 2 ms period, samples inside the timed region"}
        try:
            import subprocess

            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1, "source": "nvidia-smi after the timed region"}
        except Exception:  # noqa: BLE001
            return None


def cpu_reference(X, T, dbc, x0, steps, warmup=0, threads=None, iterations=ITERS):
    """Times the reference's CPU VBD arithmetic on this host.  Returns (value, seconds per step, info)."""
    import oracle

    kind = "reference" if oracle.have_ref() else "port"
    o = oracle.Oracle(X, T, dbc=dbc, accel=oracle.ACCEL_CHEBYSHEV, rho=RHO, kind=kind)
    # torchrun exports OMP_NUM_THREADS=1 to its children: pin the thread count explicitly
    o.set_num_threads(int(threads or os.environ.get("VBDX_REF_THREADS", 0) or os.cpu_count() or 1))
    cores = o.num_threads
    o.x = x0
    n_active = o.get("Padj").size
    for _ in range(warmup):
        o.step(DT, iterations, 1)
    t = time.perf_counter()
    for _ in range(steps):
        o.step(DT, iterations, 1)
    el = time.perf_counter() - t
    value = n_active * iterations * steps / el
    info = {"value": value, "unit": "vertex-iterations/s", "cores": cores, "kind": kind,
            "sample": f"{steps} step(s) of {iterations} iteration(s) of the same workload (Chebyshev) in {el:.1f} s; "
                      f"double precision, OpenMP over each colour, {cores} threads"}
    return value, el / steps, info


def traffic_for(key):
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        return None, None
    e = table.get("entries", {}).get(key)
    if not e:
        return None, None
    return e.get("dram_bytes_per_launch"), e.get("source")


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
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-records (N > 1: parity / strong scaling / batch; N = 1: the other colouring)")
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
        # the arm's own workload: one 58^3 slab per GPU of the domain-decomposed beam (N slabs), or one replica.  Bounded:
        # every timed step runs ITERS // slabs iterations (>= 1) of the N-slab beam, i.e. about the CPU work of one full
        # step of one slab, so that K steps + W warm-up steps end within minutes whatever N is.
        slabs = 1 if args.replicas else max(args.gpus, 1)
        X, T, dbc, x0 = workload(slabs=slabs)
        config["tets"], config["vertices"] = int(T.shape[1]), int(X.shape[1])
        it = max(1, ITERS // slabs)
        value, spstep, info = cpu_reference(X, T, dbc, x0, steps=max(steps, 1), warmup=min(warmup, 2), iterations=it)
        if it != ITERS:
            info["sample"] = f"bounded: every step runs {it} of the {ITERS} iterations; " + info["sample"]
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

        # stdout carries ONE JSON line: NCCL prints its version banner there when the first communicator is created
        # (NCCL_DEBUG=VERSION/WARN) -- send file descriptor 1 to stderr while that happens
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def make_decomposed(Xg, Tg, dbc_g, x0g):
        """Domain-decomposed integrator over all ranks, started from x0g.  The sweeps are barrier-free across the GPUs
        (DESIGN.md 5b / 6); every dependency wait carries a time-out.  Two probe steps: should any rank see a time-out, ALL
        ranks rebuild with colour barriers instead of reporting nothing."""
        from physicsbasedanimationtoolkit_b200.dist import DomainDecomposedIntegrator

        os.environ.setdefault("VBDX_DIST_TIMEOUT_S", "10")  # the ranks start together here: a peer that is 10 s late is gone
        dd = DomainDecomposedIntegrator(Xg, Tg, dbc=dbc_g, rho_chebyshev=RHO, axis=0, tile_iters=args.tile_iters)
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
        return dd

    def device_timed(vbd, nsteps, nwarm, iters=ITERS):
        """K steps back to back on a dedicated stream, CUDA events around them; returns total ms (this rank)."""
        barrier()  # domain decomposition: the ranks' kernels wait for each other, start them together
        for _ in range(nwarm):
            vbd.step(DT, iters, 1)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream = torch.cuda.Stream()  # events must sit on the stream the kernels are launched on
        vbd.use_stream(stream.cuda_stream)
        t0 = time.time()
        ev0.record(stream)
        for _ in range(nsteps):
            vbd.step_async(DT, iters, 1)
        ev1.record(stream)
        vbd.synchronize()
        barrier()
        t1 = time.time()
        vbd.use_stream(0)
        return ev0.elapsed_time(ev1), t0, t1

    def max_over_ranks(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def single_gpu_reference_run(Xg, Tg, dbc_g, x0g, colors, nsteps, time_steps=0):
        """Rank 0: the same mesh on ONE GPU from the same state; returns (positions after nsteps, ms/step or None)."""
        d = pbat.sim.vbd.Data().with_volume_mesh(Xg, Tg).with_dirichlet_vertices(dbc_g).with_chebyshev_acceleration(RHO).construct()
        assert np.array_equal(d.colors, colors)
        one = pbat.gpu.vbd.Integrator(d, device=local_rank, tile_iters=args.tile_iters)
        one.x = np.ascontiguousarray(x0g, dtype=np.float32)
        one.v = np.zeros((3, Xg.shape[1]), np.float32)
        for _ in range(nsteps):
            one.step(DT, ITERS, 1)
        x = one.x
        ms = None
        if time_steps:
            ms = float(np.median([(one.step(DT, ITERS, 1), one.info["lastStepMs"])[1] for _ in range(time_steps)]))
        del one
        return x, ms

    def parity_against_single_gpu(dd, Xg, Tg, dbc_g, x0g, nsteps=3, time_steps=0):
        """All ranks restart from x0g and take nsteps; rank 0 repeats them on one GPU; returns the record (rank 0)."""
        lp = dd.local
        dd.vbd.x = np.ascontiguousarray(x0g[:, lp.l2g], dtype=np.float32)
        dd.vbd.v = np.zeros((3, lp.l2g.size), np.float32)
        barrier()
        for _ in range(nsteps):
            dd.vbd.step(DT, ITERS, 1)
        xg = dd.gather_x()
        rec, ms1 = None, None
        if rank == 0:
            x1, ms1 = single_gpu_reference_run(Xg, Tg, dbc_g, x0g, dd.colors, nsteps, time_steps)
            diff = np.abs(xg - x1)
            rec = {"max_abs": float(diff.max()), "steps": nsteps, "vertices": int(Xg.shape[1]), "bitwise_equal": bool(np.array_equal(xg, x1)),
                   "what": "positions of the domain-decomposed run vs the same mesh, colours and start state on one GPU (rank 0)"}
        barrier()
        return rec, ms1

    decomposed = world > 1 and not args.replicas
    if decomposed:
        Xg, Tg, dbc_g, x0g = workload(seed=0, slabs=world)
        config["tets"], config["vertices"] = int(Tg.shape[1]), int(Xg.shape[1])
        dd = make_decomposed(Xg, Tg, dbc_g, x0g)
        vbd, lp = dd.vbd, dd.local
        X, T, dbc = lp.X, lp.T, np.concatenate([lp.dbc, lp.ghost_local])
        x0 = x0g[:, lp.l2g]
        nV, nT = X.shape[1], T.shape[1]
        info = vbd.info
        n_active = info["nActiveVertices"]
        n_active_job = int(Xg.shape[1] - dbc_g.size)
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
        info = vbd.info
        n_active = info["nActiveVertices"]
        n_active_job = n_active * world
    barrier_sweep = os.environ.get("VBDX_DATAFLOW") == "0"
    active = np.ones(nV, bool)
    active[dbc] = False
    kbar, nbar = mesh_stats(T, nV, active)
    B_survey = survey_bytes_per_vertex_iteration(kbar, nbar)
    fmt = format_bytes_per_vertex_iteration(info)

    x_start = np.ascontiguousarray(x0, dtype=np.float32)
    v_zero = np.zeros((3, nV), np.float32)

    # ---- device-resident leg: K steps back to back, per-step kernel time from CUDA events on the
    # launching stream (recorded inside the library around the cooperative launch)
    vbd.x = x_start
    vbd.v = v_zero
    sampler = ClockSampler(local_rank)
    launches0 = None
    if decomposed:
        vbd.dist_stats()  # reset the halo diagnostics
    barrier()
    for _ in range(warmup):
        vbd.step(DT, ITERS, 1)
    launches0 = vbd.info["kernelLaunches"]
    total_ms, t0, t1 = device_timed(vbd, steps, 0)
    clocks = sampler.stop(t0, t1)
    halo = vbd.dist_stats() if decomposed else None
    launches = vbd.info["kernelLaunches"] - launches0
    # per-launch kernel duration (events around a single launch)
    kms = []
    for _ in range(min(steps, 5)):
        vbd.step(DT, ITERS, 1)
        kms.append(vbd.info["lastStepMs"])
    kernel_ms = float(np.mean(kms))
    final_info = vbd.info
    assert np.isfinite(vbd.x).all() and final_info["nonFiniteVertices"] == 0, "non-finite positions after the timed steps"

    # ---- end-to-end leg through the public API with HOST buffers (page-locked, pbat.host.pinned_empty):
    # every step uploads the input positions, steps, and reads the result back
    vbd.x = x_start
    vbd.v = v_zero
    xin = pbat.host.pinned_empty((3, nV), np.float32)
    xout = pbat.host.pinned_empty((3, nV), np.float32)
    xin[...] = x_start
    def e2e_step(xin, xout):
        vbd.set_positions_async(xin)    # H2D of this step's input positions
        vbd.step_async(DT, ITERS, 1)
        vbd.positions_async(xout)       # D2H of the step's result
        vbd.synchronize()               # ... which the host holds when the step returns (one host round trip per step)

    for _ in range(min(warmup, 2)):
        e2e_step(xin, xout)
        xin, xout = xout, xin
    barrier()
    te = time.perf_counter()
    for _ in range(steps):
        e2e_step(xin, xout)
        xin, xout = xout, xin           # the result is the next step's input
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - te

    total_ms, e2e_ms = max_over_ranks(total_ms, e2e_s * 1e3)
    work = n_active_job * ITERS * steps
    value = work / (total_ms * 1e-3)
    e2e_value = work / (e2e_ms * 1e-3)

    # ---- N > 1: correctness against one GPU, strong scaling of configs[3], scene batches of configs[4]
    extras = {}
    if decomposed and not args.no_extras:
        rec, _ = parity_against_single_gpu(dd, Xg, Tg, dbc_g, x0g, nsteps=3)
        if rank == 0:
            extras["parity_vs_single_gpu"] = rec
        del dd, vbd
        barrier()
        # strong scaling: ONE 117^3 block over the N GPUs
        Xs, Ts, dbc_s, x0s = workload(seed=0, slabs=1, grid=STRONG_GRID)
        t_setup = time.perf_counter()
        dds = make_decomposed(Xs, Ts, dbc_s, x0s)
        t_setup = time.perf_counter() - t_setup
        dds.vbd.x = np.ascontiguousarray(x0s[:, dds.local.l2g], dtype=np.float32)
        dds.vbd.v = np.zeros((3, dds.local.l2g.size), np.float32)
        s_steps = 10
        ms_s, _, _ = device_timed(dds.vbd, s_steps, 2)
        (ms_s,) = max_over_ranks(ms_s)
        srec, ms1 = parity_against_single_gpu(dds, Xs, Ts, dbc_s, x0s, nsteps=2, time_steps=5)
        if rank == 0:
            n_act_s = int(Xs.shape[1] - dbc_s.size)
            v_s = n_act_s * ITERS * s_steps / (ms_s * 1e-3)
            sinfo = dds.vbd.info
            fmt_s = format_bytes_per_vertex_iteration(sinfo)
            peak_s = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
            extras["strong"] = {
                "workload": f"configs[3]: one {STRONG_GRID}^3-cube block, {Ts.shape[1]} tets, {Xs.shape[1]} vertices, {ITERS} iterations/step, Chebyshev {RHO}, x-slabs over {world} GPUs",
                "scaling": "strong", "n_gpus": world, "steps": s_steps, "warmup": 2, "ms_per_step": ms_s / s_steps, "value": v_s,
                "unit": "vertex-iterations/s", "single_gpu_ms_per_step": ms1, "speedup_vs_single_gpu": (ms1 / (ms_s / s_steps)) if ms1 else None,
                "format_bytes_per_vertex_iteration": fmt_s["total"],
                "roofline_frac_per_gpu": v_s * fmt_s["total"] / 1e9 / world / peak_s,
                "parity_vs_single_gpu": srec, "setup_s": t_setup,
                "timing": "CUDA events around the steps on every rank, max over ranks; the single-GPU time is the median of 5 steps of the same block on rank 0 in this same run"}
        del dds
        barrier()
    if world == 1 and not args.no_extras:
        # The same workload under the reference's other colour selection (with_vertex_coloring_strategy(LargestDegree,
        # FirstAvailable), graph/Color.h:45-135): 4 colours instead of 7 on this grid = fewer dependent tile rounds per sweep
        # (SURVEY.md 8f rank 2).  A sub-record: the headline keeps the reference's default colouring.
        data4 = (pbat.sim.vbd.Data().with_volume_mesh(X, T).with_dirichlet_vertices(dbc).with_chebyshev_acceleration(RHO)
                 .with_vertex_coloring_strategy(pbat.graph.GreedyColorOrderingStrategy.LargestDegree,
                                                pbat.graph.GreedyColorSelectionStrategy.FirstAvailable).construct())
        vbd4 = pbat.gpu.vbd.Integrator(data4, device=local_rank, tile_iters=args.tile_iters)
        vbd4.x = np.ascontiguousarray(x0, dtype=np.float32)
        c_steps = 10
        ms_c, _, _ = device_timed(vbd4, c_steps, 3)
        info4 = vbd4.info
        extras["colouring_first_available"] = {
            "what": "same workload, vertex colouring LargestDegree / FirstAvailable (a reference option) instead of the default LeastUsed",
            "colours": int(info4["nColors"]), "steps": c_steps, "ms_per_step": ms_c / c_steps,
            "value": info4["nActiveVertices"] * ITERS * c_steps / (ms_c * 1e-3), "unit": "vertex-iterations/s",
            "finite": bool(np.isfinite(vbd4.x).all())}
        del vbd4
    if world > 1 and not args.no_extras and not args.replicas:
        # scene batches (configs[4]): 512 independent 10^3-cube scenes per GPU, one persistent launch per step and GPU
        from physicsbasedanimationtoolkit_b200 import meshes

        Xb, Tb = meshes.tet_grid(10, 10, 10, 0.1)
        fixed = np.flatnonzero(Xb[2] == 0)
        datas = []
        for s in range(BATCH_SCENES_PER_GPU):
            xs = Xb + 0.002 * np.random.default_rng(rank * BATCH_SCENES_PER_GPU + s).uniform(-1, 1, Xb.shape)
            datas.append(pbat.sim.vbd.Data().with_volume_mesh(xs, Tb).with_dirichlet_vertices(fixed).construct())
        batch = pbat.gpu.vbd.BatchIntegrator(datas, device=local_rank)
        b_steps = 20
        ms_b, _, _ = device_timed(batch, b_steps, 3, iters=BATCH_ITERS)
        (ms_b,) = max_over_ranks(ms_b)
        ok_b = bool(np.isfinite(batch.x).all())
        if rank == 0:
            nact_b = batch.info["nActiveVertices"] * world
            extras["batch"] = {
                "workload": f"configs[4]: {BATCH_SCENES_PER_GPU * world} independent scenes of 10^3 cubes (5,000 tets each), {BATCH_SCENES_PER_GPU} per GPU, "
                            f"{BATCH_ITERS} iterations/step, no communication",
                "scaling": "weak", "n_gpus": world, "scenes": BATCH_SCENES_PER_GPU * world, "steps": b_steps, "ms_per_step": ms_b / b_steps,
                "value": nact_b * BATCH_ITERS * b_steps / (ms_b * 1e-3), "unit": "vertex-iterations/s",
                "scene_steps_per_s": BATCH_SCENES_PER_GPU * world * b_steps / (ms_b * 1e-3), "finite": ok_b}
        del batch
        barrier()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        bytes_per_launch = n_active * ITERS * fmt["total"]
        achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        survey_achieved = n_active * ITERS * B_survey / (kernel_ms * 1e-3) / 1e9
        tkey = f"config2/n{world}" + ("/replicas" if args.replicas else "") + ("/barrier" if barrier_sweep else "")
        traffic, traffic_src = traffic_for(tkey)
        kernel_name = "vbdx::StepKernelPipe<true,false,false,false>" if barrier_sweep else \
            "vbdx::StepKernelFlow<true,false,%s>" % ("true" if decomposed else "false")
        line = {
            "metric": "VBD vertex-iterations/sec", "value": value, "unit": "vertex-iterations/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "sweep": "barrier" if barrier_sweep else "barrier-free (dataflow-synchronised colours)",
            "steps_per_s": steps * world / (total_ms * 1e-3),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "vertex-iterations/s", "h2d_bytes_per_step": int(nV * 12),
                    "d2h_bytes_per_step": int(nV * 12), "ms_per_step": e2e_ms / steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "traffic_key": tkey,
                         "kernel": kernel_name + " (one persistent cooperative launch per step)",
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": bytes_per_launch,
                         "bytes_per_vertex_iteration": fmt["total"],
                         "bytes_model": "record format: 32 B x incident tets + 4 B x staged vertices + 16 B x tiles streamed from HBM, "
                                        "+ 96 B per-vertex state (Chebyshev); the 1-ring gather is served by the L2 and listed separately",
                         "streamed_bytes_per_vertex_iteration": fmt["stream"], "state_bytes_per_vertex_iteration": fmt["state"],
                         "l2_gather_bytes_per_vertex_iteration": fmt["l2_gather"],
                         "kbar": kbar, "nbar": nbar,
                         "frac_survey_formula": survey_achieved / peak, "survey_bytes_per_vertex_iteration": B_survey,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6.65 TB/s",
                         "streamed_record_bytes_per_launch": int(info["nRecordSlots"] * 32 * ITERS),
                         # what the DRAM counters saw (ncu, profiles/traffic.json) over the live kernel time
                         "traffic_frac": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None},
        }
        line.update(extras)
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
