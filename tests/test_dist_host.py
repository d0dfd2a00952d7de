"""CPU tests (gloo, world_size 2) of the domain-decomposition host logic: partition, local problems and
the exchange lists, including a numpy emulation of the per-colour halo exchange that must reproduce the
single-domain Gauss-Seidel sweep of a linear stand-in problem exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import dist as dd
from physicsbasedanimationtoolkit_b200 import meshes


def test_partition_and_local_problem_serial():
    X, T = meshes.tet_grid(6, 3, 3, 0.1)
    colors = pbat.graph.mesh_greedy_color(T, X.shape[1])
    owner = dd.partition_slabs(X, 3, axis=0)
    assert np.bincount(owner).tolist() == [X.shape[1] // 3 + (1 if r < X.shape[1] % 3 else 0) for r in range(3)] or \
        abs(np.bincount(owner).max() - np.bincount(owner).min()) <= 1
    seen_tets = np.zeros(T.shape[1], int)
    for r in range(3):
        lp = dd.LocalProblem(r, owner, X, T, colors)
        assert (owner[lp.l2g[:lp.n_owned]] == r).all() and (owner[lp.l2g[lp.n_owned:]] != r).all()
        assert np.array_equal(lp.l2g[lp.T], T[:, lp.tet_ids])               # same tets, local numbering
        # every tet incident to an owned vertex is local, so owned vertices see their full 1-ring
        inc = (owner[T] == r).any(axis=0)
        assert np.array_equal(np.flatnonzero(inc), lp.tet_ids)
        seen_tets[lp.tet_ids] += 1
        assert (meshes.tet_volumes(lp.X, lp.T) > 0).all()
    assert (seen_tets >= 1).all()


@pytest.mark.parametrize("nparts", [2, 3, 5, 8])
def test_recursive_coordinate_bisection(nparts):
    X, T = meshes.tet_grid(7, 6, 5, 0.1)
    colors = pbat.graph.mesh_greedy_color(T, X.shape[1])
    owner = dd.partition_rcb(X, nparts)
    counts = np.bincount(owner, minlength=nparts)
    assert counts.min() > 0 and counts.max() - counts.min() <= nparts          # balanced up to rounding at every level
    assert np.array_equal(owner, dd.partition_rcb(X, nparts))                  # deterministic
    # compact parts: every part's bounding box holds few foreign vertices compared with a random assignment
    seen = np.zeros(T.shape[1], int)
    for r in range(nparts):
        lp = dd.LocalProblem(r, owner, X, T, colors, dbc=np.flatnonzero(X[2] == 0))
        seen[lp.tet_ids] += 1
        assert lp.ghost_local.size < lp.n_owned * 4
        assert set(np.unique(lp.ghost_owner)) <= set(range(nparts)) - {r}
    assert (seen >= 1).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, T = meshes.tet_grid(8, 3, 2, 0.1)
        nV = X.shape[1]
        colors = pbat.graph.mesh_greedy_color(T, nV)
        owner = dd.partition_slabs(X, world, axis=0)
        lp = dd.LocalProblem(rank, owner, X, T, colors)
        # stand-in for the device's internal numbering: any permutation of the local vertices
        internal = np.random.default_rng(rank).permutation(lp.l2g.size)
        sl, sp, sr = dd.exchange_lists(lp, internal[lp.ghost_local], world, dd.torch_all_to_all)
        assert (sl < lp.n_owned).all() and (sp != rank).all()
        # emulate the halo exchange with gloo: a Jacobi-free coloured Gauss-Seidel smoothing of a scalar field
        # (x_i <- mean of its 1-ring), colour by colour, ghosts refreshed after every colour
        nbrs = [set() for _ in range(lp.l2g.size)]
        for e in range(lp.T.shape[1]):
            for a in lp.T[:, e]:
                nbrs[a].update(int(b) for b in lp.T[:, e] if b != a)
        f = np.sin(7.0 * lp.X[0]) + lp.X[1]                      # same global initial field on every rank
        slot_of_internal = np.empty(lp.l2g.size, np.int64)
        slot_of_internal[internal] = np.arange(lp.l2g.size)      # receiver: internal slot -> local vertex
        for sweep in range(3):
            for c in range(int(colors.max()) + 1):
                mine = [i for i in range(lp.n_owned) if lp.colors[i] == c]
                new = {i: np.mean([f[j] for j in nbrs[i]]) for i in mine}
                for i, val in new.items():
                    f[i] = val
                # push the updated owned values of colour c to the peers that hold them as ghosts
                msgs = []
                for r in range(world):
                    sel = (sp == r) & (lp.colors[sl] == c)
                    msgs.append(np.concatenate([sr[sel].astype(np.float64), f[sl[sel]]]))
                got = _a2a_f64(msgs)
                for r in range(world):
                    m = got[r]
                    k = m.size // 2
                    f[slot_of_internal[m[:k].astype(np.int64)]] = m[k:]
        res = np.zeros(nV)
        res[lp.l2g[:lp.n_owned]] = f[:lp.n_owned]
        t = torch.as_tensor(res)
        dist.all_reduce(t)
        if rank == 0:
            out.put(t.numpy())
    finally:
        dist.destroy_process_group()


def _a2a_f64(arrays):
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = torch.tensor([len(a) for a in arrays], dtype=torch.int64)
    all_sizes = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = torch.stack(all_sizes).numpy()
    width = max(int(all_sizes.max()), 1)
    payload = torch.zeros((world, width), dtype=torch.float64)
    for r, a in enumerate(arrays):
        payload[r, :len(a)] = torch.as_tensor(a)
    gathered = [torch.zeros_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload)
    return [gathered[src][rank, :all_sizes[src, rank]].numpy() for src in range(world)]


@pytest.mark.timeout(120)
def test_exchange_lists_reproduce_the_global_sweep_world2():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    # single-domain reference of the same coloured Gauss-Seidel smoothing
    X, T = meshes.tet_grid(8, 3, 2, 0.1)
    nV = X.shape[1]
    colors = pbat.graph.mesh_greedy_color(T, nV)
    nbrs = [set() for _ in range(nV)]
    for e in range(T.shape[1]):
        for a in T[:, e]:
            nbrs[a].update(int(b) for b in T[:, e] if b != a)
    f = np.sin(7.0 * X[0]) + X[1]
    for sweep in range(3):
        for c in range(int(colors.max()) + 1):
            new = {i: np.mean([f[j] for j in nbrs[i]]) for i in range(nV) if colors[i] == c}
            for i, val in new.items():
                f[i] = val
    assert np.allclose(got, f, rtol=0, atol=1e-13)
