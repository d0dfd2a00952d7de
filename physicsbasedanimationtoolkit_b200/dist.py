"""Domain decomposition of one mesh across the GPUs of a node (first row).

The reference has no multi-GPU path; this is new.  Ownership is by vertex.  Every rank holds the
vertices it owns, one layer of *ghost* vertices (every vertex that shares a tet with an owned vertex
but is owned elsewhere) and every tet incident to an owned vertex.  Ghosts are never swept locally:
when a boundary vertex is swept its owner pushes the new position straight into the ghost slots of the
peers (NVLink peer-to-peer stores inside the persistent step kernel, tagged with the number of the
write so that readers can wait on the datum itself).  Colours are computed once on the global mesh so
that colour c means the same on every rank.

Everything in this module is host logic (numpy + ``torch.distributed`` collectives for the plumbing):
partitioning, local problems, and the exchange lists.  It runs on CPU with the ``gloo`` backend, which
is how the tests cover it; the device side is ``vbdx_dist_*`` in ``include/vbdx.h``.
"""
from __future__ import annotations

import numpy as np


def partition_slabs(X, nparts: int, axis: int = 0):
    """Owner rank of every vertex: ``nparts`` slabs along ``axis`` with (almost) equal vertex counts;
    ties on the coordinate are broken by vertex id so that the result is deterministic."""
    n = X.shape[1]
    order = np.lexsort((np.arange(n), X[axis]))
    owner = np.empty(n, dtype=np.int64)
    bounds = (np.arange(nparts + 1) * n) // nparts
    for r in range(nparts):
        owner[order[bounds[r]:bounds[r + 1]]] = r
    return owner


def partition_rcb(X, nparts: int):
    """Owner rank of every vertex by recursive coordinate bisection (general meshes): the
    current vertex set is split along the longest axis of its bounding box into two sets whose sizes are proportional
    to the numbers of ranks they will hold, recursively; ties on the coordinate are broken by vertex id.  A rank may
    then have more than two neighbours (the device side allows up to 8 ranks in all)."""
    n = X.shape[1]
    owner = np.empty(n, dtype=np.int64)

    def split(idx, first, parts):
        if parts == 1:
            owner[idx] = first
            return
        ext = X[:, idx].max(axis=1) - X[:, idx].min(axis=1)
        ax = int(np.argmax(ext))
        order = idx[np.lexsort((idx, X[ax, idx]))]
        left = parts // 2
        cut = (order.size * left) // parts
        split(order[:cut], first, left)
        split(order[cut:], first + left, parts - left)

    split(np.arange(n), 0, int(nparts))
    return owner


class LocalProblem:
    """What one rank simulates: global ids of its local vertices (owned first, ascending; then the
    constrained vertices of other ranks its tets touch; then ghosts, ascending), the local tets in local
    numbering, and which local vertices are ghosts."""

    def __init__(self, rank, owner, X, T, colors, dbc=None, v=None):
        owner = np.asarray(owner)
        mine = owner == rank
        tet_mask = mine[T].any(axis=0)                      # tets incident to an owned vertex
        Tl = T[:, tet_mask]
        touched = np.zeros(owner.size, dtype=bool)
        touched[Tl.reshape(-1)] = True
        isd = np.zeros(owner.size, dtype=bool)
        if dbc is not None and len(dbc):
            isd[np.asarray(dbc)] = True
        owned = np.flatnonzero(mine)
        # constrained vertices of other ranks never move: they are plain Dirichlet vertices here, not ghosts
        fixed = np.flatnonzero(touched & ~mine & isd)
        ghosts = np.flatnonzero(touched & ~mine & ~isd)
        self.rank = rank
        self.l2g = np.concatenate([owned, fixed, ghosts])
        self.n_owned = owned.size
        g2l = np.full(owner.size, -1, dtype=np.int64)
        g2l[self.l2g] = np.arange(self.l2g.size)
        self.g2l = g2l
        self.T = np.ascontiguousarray(g2l[Tl])
        self.tet_ids = np.flatnonzero(tet_mask)
        self.X = np.ascontiguousarray(X[:, self.l2g])
        self.colors = np.ascontiguousarray(np.asarray(colors)[self.l2g])
        self.ghost_local = np.arange(self.n_owned + fixed.size, self.l2g.size)
        self.ghost_owner = owner[ghosts]
        # local Dirichlet set = the real constraints among the local vertices; ghosts are never swept either
        self.dbc = np.flatnonzero(isd[self.l2g])
        self.v = None if v is None else np.ascontiguousarray(v[:, self.l2g])


def exchange_lists(local: LocalProblem, ghost_internal_ids, world: int, all_to_all):
    """Who sends what to whom.

    ``ghost_internal_ids[k]`` is this rank's device-internal id of its k-th ghost (the slot a peer must
    write).  ``all_to_all(list_of_arrays) -> list_of_arrays`` exchanges one int64 array with every rank.
    Returns ``(send_local, send_peer, send_remote)``: for every owned vertex that is a ghost on some peer,
    the local (caller-order) id, the peer rank and the peer's internal id of that ghost.
    """
    out = []
    for r in range(world):
        sel = local.ghost_owner == r
        # pairs (global id, my internal slot) for the ghosts owned by r
        out.append(np.stack([local.l2g[local.ghost_local[sel]], np.asarray(ghost_internal_ids)[sel]]).astype(np.int64).reshape(-1))
    got = all_to_all(out)
    send_local, send_peer, send_remote = [], [], []
    for r in range(world):
        pairs = np.asarray(got[r], dtype=np.int64).reshape(2, -1)
        if pairs.shape[1] == 0:
            continue
        loc = local.g2l[pairs[0]]
        assert (loc >= 0).all() and (loc < local.n_owned).all(), "peer asked for a vertex this rank does not own"
        send_local.append(loc)
        send_peer.append(np.full(loc.size, r, dtype=np.int64))
        send_remote.append(pairs[1])
    if not send_local:
        z = np.zeros(0, dtype=np.int64)
        return z, z.copy(), z.copy()
    return np.concatenate(send_local), np.concatenate(send_peer), np.concatenate(send_remote)


def torch_all_to_all(arrays):
    """``all_to_all`` of variable-length int64 arrays over the default torch.distributed group (works with
    gloo and nccl: sizes first, then padded payloads through all_gather)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    sizes = torch.tensor([len(a) for a in arrays], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(world, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = torch.stack(all_sizes).cpu().numpy()          # [src, dst]
    width = int(all_sizes.max()) if all_sizes.size else 0
    payload = torch.zeros((world, max(width, 1)), dtype=torch.int64, device=dev)
    for r, a in enumerate(arrays):
        if len(a):
            payload[r, :len(a)] = torch.as_tensor(np.asarray(a, dtype=np.int64), device=dev)
    gathered = [torch.zeros_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload)
    return [gathered[src][rank, :all_sizes[src, rank]].cpu().numpy() for src in range(world)]


class DomainDecomposedIntegrator:
    """One mesh across the GPUs of a node: one process per GPU (``torch.distributed``, NCCL for the
    plumbing), every process builds the same global problem description, keeps its slab, and steps in
    lock-step with its peers.  ``x``/``v`` return the *owned* part; ``gather_x()`` assembles the global
    array on every rank."""

    def __init__(self, X, T, *, dbc=None, v=None, rho_chebyshev=None, colors=None, axis=0, partition="slabs", **tuning):
        import torch
        import torch.distributed as dist

        from . import graph
        from .gpu.vbd import Integrator
        from .sim.vbd import Data

        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        nV = X.shape[1]
        if colors is None:
            colors = graph.mesh_greedy_color(T, nV)         # global colouring: colour c means the same everywhere
        # partition: "slabs" along `axis` (the default, at most two neighbours per GPU), "rcb" (recursive coordinate
        # bisection, general shapes), or an explicit owner rank per vertex
        if isinstance(partition, str):
            if partition not in ("slabs", "rcb"):
                raise ValueError("partition must be 'slabs', 'rcb' or an array of owner ranks")
            owner = partition_slabs(X, self.world, axis) if partition == "slabs" else partition_rcb(X, self.world)
        else:
            owner = np.asarray(partition, dtype=np.int64)
            if owner.shape != (nV,) or owner.min() < 0 or owner.max() >= self.world:
                raise ValueError("an explicit partition needs one owner rank in [0, world) per vertex")
        self.colors = np.asarray(colors)
        self.local = lp = LocalProblem(self.rank, owner, X, T, colors, dbc, v)
        self.nV_global = nV
        data = Data().with_volume_mesh(lp.X, lp.T)
        if lp.dbc.size:
            data = data.with_dirichlet_vertices(lp.dbc)
        if lp.v is not None:
            data = data.with_velocity(lp.v)
        if rho_chebyshev:
            data = data.with_chebyshev_acceleration(rho_chebyshev)
        data = data.construct()
        data.colors = lp.colors                              # not the local greedy colouring
        # every rank sweeps the colours of the WHOLE mesh (a slab may lack the last ones: the epochs would drift apart)
        self.vbd = Integrator(data, ghosts=lp.ghost_local, kernel_variant=3, n_colors=int(np.max(colors)) + 1, **tuning)
        if self.world > 1:
            ids = self.vbd.internal_ids()
            sl, sp, sr = exchange_lists(lp, ids[lp.ghost_local], self.world, torch_all_to_all)
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
            mine = torch.as_tensor(self.vbd.ipc_handles(), device=dev)
            handles = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(handles, mine)
            nv = torch.tensor([lp.l2g.size, lp.ghost_local.size, int(self.vbd.info["nColors"])], dtype=torch.int64, device=dev)
            nvs = [torch.zeros_like(nv) for _ in range(self.world)]
            dist.all_gather(nvs, nv)
            nvs = torch.stack(nvs).cpu().numpy()
            # barrier sweeps count epochs per colour: every rank must sweep the same number of colours (empty ones included)
            if len(set(nvs[:, 2].tolist())) != 1:
                raise RuntimeError(f"ranks disagree on the number of colours: {nvs[:, 2].tolist()}")
            self.vbd.dist_connect(self.rank, self.world, torch.stack(handles).cpu().numpy(),
                                  nvs[:, 0], nvs[:, 1], sl, sp, sr,
                                  sum(1 << int(r) for r in np.unique(lp.ghost_owner)))
            dist.barrier()
        self.n_send = 0 if self.world == 1 else int(sl.size)

    def step(self, dt=0.01, iterations=20, substeps=1):
        self.vbd.step(dt, iterations, substeps)

    @property
    def x_owned(self):
        return self.vbd.x[:, :self.local.n_owned]

    def gather_x(self):
        import torch
        import torch.distributed as dist

        out = np.zeros((3, self.nV_global), np.float32)
        out[:, self.local.l2g[:self.local.n_owned]] = self.x_owned
        if self.world > 1:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
            t = torch.as_tensor(out, device=dev)
            dist.all_reduce(t)                               # owned sets are disjoint: sum == assembly
            out = t.cpu().numpy()
        return out
