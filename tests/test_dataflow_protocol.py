"""CPU model check of the barrier-free sweep against the planner's real output.

The kernels drop the colour barrier: every position carries the number of its write, and a tile may run once every entry of
its ring list carries the number it expects -- `tagLow + 1` for entries without the previous-iterate flag (neighbours of a
lower colour), `tagLow` for flagged entries (neighbours of a higher colour, the tile's own vertices); constrained vertices
and padding entries are not checked (csrc/step_kernel_pipe.cuh: AwaitTags).  Here warps are scheduled by an adversarial
random scheduler over the tiles, ring lists, flags and padding that `BuildPlan` (csrc/plan.cpp) produced, and every read is
checked:

  * progress: until all tiles are done, some warp can always run (no deadlock);
  * no overwrite hazard: when a tile's dependencies are met, no entry carries a NEWER write than expected (the poll waits
    for equality, so a newer value would also mean a hang on the device);
  * the values read are exactly the ones the barrier schedule reads (Gauss-Seidel order): lower colours from this sweep,
    higher colours from the previous one.

Reads happen when the dependencies are met, writes an arbitrary time later (other warps run in between), as on the GPU.
"""
import ctypes as C

import numpy as np
import pytest

import physicsbasedanimationtoolkit_b200 as pbat
from physicsbasedanimationtoolkit_b200 import _lib, meshes

PREV = 0x80000000


def plan_of(X, T, colors, constrained, tile_iters=0):
    L = _lib.lib()
    nV, nT = X.shape[1], T.shape[1]
    Xc = np.ascontiguousarray(X.T, np.float64)
    Ec = np.ascontiguousarray(T.T, np.int64)
    col = np.ascontiguousarray(colors, np.int64)
    con = np.ascontiguousarray(constrained, np.uint8)
    h = C.c_void_p()
    _lib.check(L.vbdx_debug_plan_create(nV, nT, Ec.ctypes.data, col.ctypes.data, con.ctypes.data, Xc.ctypes.data, tile_iters, C.byref(h)))
    try:
        sizes = np.zeros(5, np.int64)
        _lib.check(L.vbdx_debug_plan_get(h, 0, sizes.ctypes.data))
        n_tiles, n_ids, n_colors, n_active, ghost_begin = (int(v) for v in sizes)
        tiles = np.zeros((n_tiles, 4), np.uint32)
        ids = np.zeros(n_ids, np.uint32)
        ctb = np.zeros(n_colors + 1, np.uint32)
        new2old = np.zeros(nV, np.int32)
        for what, a in ((1, tiles), (2, ids), (3, ctb), (4, new2old)):
            _lib.check(L.vbdx_debug_plan_get(h, what, a.ctypes.data))
    finally:
        L.vbdx_debug_plan_destroy(h)
    return dict(tiles=tiles, ids=ids, ctb=ctb, new2old=new2old, n_active=n_active, n_colors=n_colors)


def simulate(plan, colors, iterations, n_warps, rng, neighbours):
    tiles, ids, ctb = plan["tiles"], plan["ids"], plan["ctb"]
    n_active, n_colors, new2old = plan["n_active"], plan["n_colors"], plan["new2old"]
    icolor = np.asarray(colors)[new2old]                      # colour by internal id
    nV = new2old.size
    T0 = 5
    tagQ = np.full(nV, T0, np.int64)                          # pre-step: every vertex carries T0 in both buffers
    tagP = np.full(nV, T0, np.int64)
    # per-warp programme: (sweep k, tile) in (k, colour) order, tiles of a colour dealt round-robin
    prog = [[] for _ in range(n_warps)]
    for k in range(iterations):
        for c in range(n_colors):
            for j, t in enumerate(range(int(ctb[c]), int(ctb[c + 1]))):
                prog[j % n_warps].append((k, t))
    pc = [0] * n_warps
    pending = [None] * n_warps                                # a tile that has read but not written yet
    remaining = sum(len(p) for p in prog)
    reads_checked = 0
    while remaining:
        order = rng.permutation(n_warps)
        progressed = False
        for w in order:
            if pending[w] is not None:
                if rng.random() < 0.5:                        # the write lands some time after the read
                    k, t = pending[w]
                    vbase, meta = int(tiles[t, 1]), int(tiles[t, 2])
                    nverts = (meta >> 3) & 63
                    tagQ[vbase:vbase + nverts] = T0 + k + 1
                    tagP[vbase:vbase + nverts] = T0 + k + 1
                    pending[w] = None
                    pc[w] += 1
                    remaining -= 1
                progressed = True
                continue
            if pc[w] >= len(prog[w]):
                continue
            k, t = prog[w][pc[w]]
            vbase, meta, ring_start = int(tiles[t, 1]), int(tiles[t, 2]), int(tiles[t, 3])
            chunks = (meta >> 9) & 63
            nverts = (meta >> 3) & 63
            entries = ids[ring_start:ring_start + 32 * chunks]
            base = (entries & ~np.uint32(PREV)).astype(np.int64)
            prev = (entries & np.uint32(PREV)) != 0
            checked = (base < n_active) & ~((base == vbase) & ~prev)
            expect = T0 + k + np.where(prev, 0, 1)
            have = np.where(prev, tagP[base], tagQ[base])
            newer = checked & (have > expect)
            assert not newer.any(), f"overwrite hazard: tile {t} sweep {k} would read a newer write of vertex {base[newer][0]}"
            if (checked & (have != expect)).any():
                continue                                       # keeps polling
            # dependencies met: this is what the barrier schedule reads
            my_color = icolor[vbase]
            nb = base[checked]
            lower = icolor[nb] < my_color
            assert np.array_equal(prev[checked], ~lower | (nb >= vbase) & (nb < vbase + nverts)), "previous-iterate flags"
            # every swept neighbour of every vertex of the tile is listed (else it could be overwritten unnoticed)
            for v in range(vbase, vbase + nverts):
                need = neighbours[v]
                assert np.isin(need[need < n_active], nb).all(), f"tile {t}: a swept neighbour of vertex {v} is missing from its ring list"
            reads_checked += int(checked.sum())
            pending[w] = (k, t)
            progressed = True
        assert progressed, "deadlock: no warp can proceed"
    assert (tagQ[:n_active] == T0 + iterations).all()
    return reads_checked


def ring_neighbours(T, old2new, nV):
    nb = [set() for _ in range(nV)]
    for tet in T.T:
        for a in tet:
            for b in tet:
                if a != b:
                    nb[old2new[a]].add(old2new[b])
    return [np.fromiter(s, dtype=np.int64, count=len(s)) for s in nb]


def case(X, T, dbc, tile_iters, n_warps, seed, iterations=3):
    nV = X.shape[1]
    colors = pbat.graph.mesh_greedy_color(T, nV)
    constrained = np.zeros(nV, np.uint8)
    constrained[dbc] = 1
    plan = plan_of(X, T, colors, constrained, tile_iters)
    old2new = np.empty(nV, np.int64)
    old2new[plan["new2old"]] = np.arange(nV)
    nbs = ring_neighbours(T, old2new, nV)
    return simulate(plan, colors, iterations, n_warps, np.random.default_rng(seed), nbs)


@pytest.mark.parametrize("tile_iters,n_warps", [(0, 3), (0, 64), (1, 7), (3, 1000)])
def test_protocol_on_a_grid(tile_iters, n_warps):
    X, T = meshes.tet_grid(6, 5, 4, 0.1)
    assert case(X, T, np.flatnonzero(X[0] == 0), tile_iters, n_warps, seed=tile_iters + n_warps) > 1000


def test_protocol_with_extreme_valence_and_ragged_components():
    from test_gpu_edge_cases import icosphere_star

    Xs, Ts = icosphere_star()                                   # a vertex with 80 incident tets: rings of several chunks
    Xa, Ta = meshes.tet_grid(3, 3, 2, 0.2, origin=(2.0, 0.0, 0.0))
    Xb = np.array([[5., 5.3, 5., 5.], [0., 0., 0.3, 0.], [0., 0., 0., 0.3]])
    Tb = np.array([[0], [1], [2], [3]], dtype=np.int64)
    X = np.concatenate([Xs, Xa, Xb], axis=1)
    T = np.concatenate([Ts, Ta + Xs.shape[1], Tb + Xs.shape[1] + Xa.shape[1]], axis=1)
    dbc = np.flatnonzero(X[2] > 0.25)
    for tile_iters, n_warps, seed in ((0, 5, 1), (2, 2, 2), (8, 40, 3)):
        assert case(X, T, dbc, tile_iters, n_warps, seed) > 500


def test_protocol_without_constraints_and_single_warp():
    X, T = meshes.tet_grid(3, 3, 3, 0.1)
    assert case(X, T, np.zeros(0, int), 0, 1, seed=9, iterations=4) > 500      # one warp: pure programme order


# ---------------------------------------------------------------------------------------------------------------------
# several GPUs: the same protocol across the halo.  A ghost exists twice on the reading GPU (one copy per
# parity of the write number); the owner's store of write t goes into copy t & 1 and may take arbitrarily long to arrive.
# ---------------------------------------------------------------------------------------------------------------------
def simulate_ranks(X, T, dbc, world, iterations, n_warps, rng, tile_iters=0, partition="slabs"):
    from physicsbasedanimationtoolkit_b200.dist import LocalProblem, partition_rcb, partition_slabs

    nV = X.shape[1]
    colors = pbat.graph.mesh_greedy_color(T, nV)
    owner = partition_slabs(X, world, 0) if partition == "slabs" else partition_rcb(X, world)
    T0 = 6
    ranks = []
    for r in range(world):
        lp = LocalProblem(r, owner, X, T, colors, dbc)
        con = np.zeros(lp.l2g.size, np.uint8)
        con[lp.dbc] = 1
        con[lp.ghost_local] = 2
        plan = plan_of(lp.X, lp.T, lp.colors, con, tile_iters)
        n_loc = lp.l2g.size
        glob_of_internal = lp.l2g[plan["new2old"]]
        ghost_begin = n_loc - lp.ghost_local.size
        prog = [[] for _ in range(n_warps)]
        for k in range(iterations):
            for c in range(plan["n_colors"]):
                for j, t in enumerate(range(int(plan["ctb"][c]), int(plan["ctb"][c + 1]))):
                    prog[j % n_warps].append((k, t))
        ranks.append(dict(lp=lp, plan=plan, glob=glob_of_internal, ghost_begin=ghost_begin, prog=prog, pc=[0] * n_warps,
                          pending=[None] * n_warps, tagQ=np.full(n_loc, T0, np.int64), tagP=np.full(n_loc, T0, np.int64),
                          gQ=np.full((2, n_loc), -1, np.int64), gP=np.full((2, n_loc), -1, np.int64)))
    # the pre-step pushed every ghost with T0 and a barrier followed
    for R in ranks:
        R["gQ"][T0 & 1, R["ghost_begin"]:] = T0
        R["gP"][T0 & 1, R["ghost_begin"]:] = T0
    # where an owned vertex is a ghost: global id -> [(rank, internal ghost id)]
    holders = {}
    for r, R in enumerate(ranks):
        for i in range(R["ghost_begin"], R["glob"].size):
            holders.setdefault(int(R["glob"][i]), []).append((r, i))
    wire = []                                                     # stores under way: (rank, ghost, tag)
    remaining = sum(len(p) for R in ranks for p in R["prog"])
    checked_ghost_reads = 0
    while remaining:
        progressed = False
        # some stores arrive, in any order
        rng.shuffle(wire)
        while wire and rng.random() < 0.6:
            r, i, tag = wire.pop()
            slotQ, slotP = ranks[r]["gQ"][tag & 1], ranks[r]["gP"][tag & 1]
            assert slotQ[i] < tag, "a store overtook a later store to the same ghost copy"
            slotQ[i] = slotP[i] = tag
            progressed = True
        for r in rng.permutation(world):
            R = ranks[r]
            plan = R["plan"]
            tiles, ids, n_active = plan["tiles"], plan["ids"], plan["n_active"]
            for w in rng.permutation(n_warps):
                if R["pending"][w] is not None:
                    if rng.random() < 0.5:
                        k, t = R["pending"][w]
                        vbase, nverts = int(tiles[t, 1]), (int(tiles[t, 2]) >> 3) & 63
                        tag = T0 + k + 1
                        R["tagQ"][vbase:vbase + nverts] = tag
                        R["tagP"][vbase:vbase + nverts] = tag
                        for v in range(vbase, vbase + nverts):
                            for dest in holders.get(int(R["glob"][v]), ()):
                                wire.append((dest[0], dest[1], tag))
                        R["pending"][w] = None
                        R["pc"][w] += 1
                        remaining -= 1
                    progressed = True
                    continue
                if R["pc"][w] >= len(R["prog"][w]):
                    continue
                k, t = R["prog"][w][R["pc"][w]]
                vbase, meta, ring_start = int(tiles[t, 1]), int(tiles[t, 2]), int(tiles[t, 3])
                entries = ids[ring_start:ring_start + 32 * ((meta >> 9) & 63)]
                base = (entries & ~np.uint32(PREV)).astype(np.int64)
                prev = (entries & np.uint32(PREV)) != 0
                expect = T0 + k + np.where(prev, 0, 1)
                local = (base < n_active) & ~((base == vbase) & ~prev)
                ghost = base >= R["ghost_begin"]
                have = np.where(prev, R["tagP"][base], R["tagQ"][base])
                par = expect & 1
                have_g = np.where(prev, R["gP"][par, base], R["gQ"][par, base])
                have = np.where(ghost, have_g, have)
                check = local | ghost
                assert not (check & (have > expect)).any(), f"rank {r} tile {t} sweep {k}: a value was overwritten before it was read"
                if (check & (have != expect)).any():
                    continue
                checked_ghost_reads += int(ghost.sum())
                R["pending"][w] = (k, t)
                progressed = True
        assert progressed or wire, "deadlock: no warp on any GPU can proceed and nothing is under way"
    for R in ranks:
        assert (R["tagQ"][:R["plan"]["n_active"]] == T0 + iterations).all()
    return checked_ghost_reads


@pytest.mark.parametrize("world,n_warps", [(2, 4), (3, 2), (4, 16), (8, 3)])
def test_protocol_across_gpus(world, n_warps):
    X, T = meshes.tet_grid(3 * world, 4, 3, 0.1)
    dbc = np.flatnonzero(X[2] == 0)                              # a constrained face crossing every interface
    assert simulate_ranks(X, T, dbc, world, 4, n_warps, np.random.default_rng(world * 100 + n_warps)) > 200


@pytest.mark.parametrize("world", [4, 7, 8])
def test_protocol_across_gpus_with_many_neighbours(world):
    """Recursive coordinate bisection of a cube: GPUs with up to world - 1 neighbours."""
    X, T = meshes.tet_grid(6, 6, 6, 0.1)
    dbc = np.flatnonzero(X[2] == 0)
    assert simulate_ranks(X, T, dbc, world, 3, 5, np.random.default_rng(world), partition="rcb") > 500


# ---------------------------------------------------------------------------------------------------------------------
# contact: a triangle corner is a vertex of ANOTHER body -- no ring dependency ties its tile to the reader's.
# Every write also goes to the history of the vertex' last four writes (slot = write number & 3); a reader polls the slot of
# the write it needs; no warp starts sweep k before every warp with tiles has finished sweep k - 2 (csrc/step_kernel_flow.cuh).
# ---------------------------------------------------------------------------------------------------------------------
def simulate_contact(plan, colors, iterations, n_warps, rng, contacts, lag_bound=True, favoured=None, warp_of=None):
    """`contacts`: internal vertex -> array of corner vertices (internal ids) its contact term reads.  Ring reads happen when
    the ring dependencies are met, contact reads some time later, the write later still.  `favoured`: warps the scheduler
    prefers (the others only run when no favoured warp can) -- lets one body run as far ahead as the protocol allows.
    `warp_of(colour, index within the colour, tile)`: which warp runs a tile (default: dealt round-robin like the kernel's
    schedule; the protocol must hold for ANY assignment in which a warp runs its tiles in (sweep, colour) order)."""
    tiles, ids, ctb = plan["tiles"], plan["ids"], plan["ctb"]
    n_active, n_colors, new2old = plan["n_active"], plan["n_colors"], plan["new2old"]
    icolor = np.asarray(colors)[new2old]
    nV = new2old.size
    T0 = 9
    tagQ = np.full(nV, T0, np.int64)
    tagP = np.full(nV, T0, np.int64)
    hist = np.full((4, nV), -1, np.int64)
    hist[T0 & 3] = T0                                          # the pre-step primes the history
    prog = [[] for _ in range(n_warps)]
    for k in range(iterations):
        for c in range(n_colors):
            for j, t in enumerate(range(int(ctb[c]), int(ctb[c + 1]))):
                prog[warp_of(c, j, t) if warp_of else j % n_warps].append((k, t))
    active_warps = sum(1 for p in prog if p)
    pc = [0] * n_warps
    stage = [0] * n_warps                                      # 0: waiting for the ring, 1: ring read, 2: contacts read
    sweeps_done = [0] * iterations                             # per sweep: warps that have finished it (one add per warp and sweep)
    remaining = sum(len(p) for p in prog)
    contact_reads = 0
    max_lead = 0
    while remaining:
        order = list(rng.permutation(n_warps))
        if favoured is not None:
            order.sort(key=lambda w: w not in favoured)
        progressed = False
        for w in order:
            if favoured is not None and w not in favoured and progressed:
                break                                          # the others run only when no favoured warp can
            if pc[w] >= len(prog[w]):
                continue
            k, t = prog[w][pc[w]]
            vbase, meta, ring_start = int(tiles[t, 1]), int(tiles[t, 2]), int(tiles[t, 3])
            nverts = (meta >> 3) & 63
            if stage[w] == 0:
                first_of_sweep = pc[w] == 0 or prog[w][pc[w] - 1][0] != k
                if lag_bound and first_of_sweep and k >= 2 and sweeps_done[k - 2] < active_warps:
                    continue                                   # polls the counter
                entries = ids[ring_start:ring_start + 32 * ((meta >> 9) & 63)]
                base = (entries & ~np.uint32(PREV)).astype(np.int64)
                prev = (entries & np.uint32(PREV)) != 0
                checked = (base < n_active) & ~((base == vbase) & ~prev)
                expect = T0 + k + np.where(prev, 0, 1)
                have = np.where(prev, tagP[base], tagQ[base])
                assert not (checked & (have > expect)).any(), "ring overwrite hazard"
                if (checked & (have != expect)).any():
                    continue
                stage[w] = 1
                progressed = True
                if favoured is None or w not in favoured:
                    continue                                   # (favoured warps go on at once: as far ahead as possible)
            if stage[w] == 1:
                if favoured is None and rng.random() < 0.5:
                    progressed = True
                    continue                                   # the contact reads come some time after the ring reads
                ready = True
                for v in range(vbase, vbase + nverts):
                    for j in contacts.get(v, ()):
                        if j >= n_active:
                            continue                           # never swept: read in place
                        lower = icolor[j] < icolor[v]
                        want = T0 + k + (1 if lower else 0)
                        have = hist[want & 3, j]
                        assert have <= want, (f"history overwritten: vertex {v} (sweep {k}) needs write {want} of vertex {j}, "
                                              f"its slot already holds write {have}")
                        if have != want:
                            ready = False
                        else:
                            contact_reads += 1
                if not ready:
                    continue                                   # keeps polling (AwaitHist)
                stage[w] = 2
                progressed = True
                if favoured is None or w not in favoured:
                    continue
            # stage 2: the write
            if favoured is None and rng.random() < 0.5:
                progressed = True
                continue
            tag = T0 + k + 1
            tagQ[vbase:vbase + nverts] = tag
            tagP[vbase:vbase + nverts] = tag
            hist[tag & 3, vbase:vbase + nverts] = tag
            stage[w] = 0
            pc[w] += 1
            remaining -= 1
            progressed = True
            if pc[w] == len(prog[w]) or prog[w][pc[w]][0] != k:
                sweeps_done[k] += 1                            # this warp's last tile of sweep k
            lo = min((prog[x][pc[x]][0] if pc[x] < len(prog[x]) else iterations) for x in range(n_warps) if prog[x])
            max_lead = max(max_lead, k - lo)
        assert progressed, "deadlock: no warp can proceed"
    assert (tagQ[:n_active] == T0 + iterations).all()
    return contact_reads, max_lead


def contact_scene(seed, one_way=False, tile_iters=0):
    """Two grids that share no tet; random contact lists (<= 8 triangles = 24 corners) between them."""
    Xa, Ta = meshes.tet_grid(4, 4, 2, 0.1)
    Xb, Tb = meshes.tet_grid(4, 3, 2, 0.1, origin=(0.0, 0.0, 0.25))
    X = np.concatenate([Xa, Xb], axis=1)
    T = np.concatenate([Ta, Tb + Xa.shape[1]], axis=1)
    nA, nV = Xa.shape[1], X.shape[1]
    colors = pbat.graph.mesh_greedy_color(T, nV)
    constrained = np.zeros(nV, np.uint8)
    constrained[np.flatnonzero(X[2] == 0)] = 1
    plan = plan_of(X, T, colors, constrained, tile_iters)
    old2new = np.empty(nV, np.int64)
    old2new[plan["new2old"]] = np.arange(nV)
    rng = np.random.default_rng(seed)
    contacts = {}
    body = np.arange(nV) >= nA
    for v in rng.choice(nV, 30, replace=False):
        if one_way and body[v]:
            continue                                           # only body A reads body B
        other = np.flatnonzero(body != body[v])
        iv = int(old2new[v])
        if iv < plan["n_active"]:
            contacts[iv] = old2new[rng.choice(other, 3 * int(rng.integers(1, 9)))]
    tile_body = []                                             # None: the tile holds vertices of both bodies
    for t in plan["tiles"]:
        b = body[plan["new2old"][int(t[1]):int(t[1]) + ((int(t[2]) >> 3) & 63)]]
        tile_body.append(bool(b[0]) if (b == b[0]).all() else None)
    return plan, colors, contacts, tile_body


@pytest.mark.parametrize("n_warps,seed", [(2, 1), (7, 2), (64, 3), (500, 4)])
def test_contact_history_protocol(n_warps, seed):
    plan, colors, contacts, _ = contact_scene(seed)
    reads, lead = simulate_contact(plan, colors, 7, n_warps, np.random.default_rng(seed), contacts)
    assert reads > 500 and lead <= 1


def test_contact_history_needs_the_sweep_lag_bound():
    """Two bodies that share no tile (a hand-made plan: two paths of six vertices, two colours, one vertex per tile and warp);
    A's contact lists read B, B reads nothing, and the scheduler lets B's warps go whenever they can.  With the bound B is
    never more than a sweep ahead of the slowest warp and every history read finds its write; without it B finishes all
    its sweeps before A starts, laps the four slots, and the model check reports the overwrite -- the bound is what makes
    the history safe, not luck."""
    # internal ids: colour 0 = A0 A2 A4 B0 B2 B4 (0..5), colour 1 = A1 A3 A5 B1 B3 B5 (6..11); position along the path:
    pos = {0: ("A", 0), 1: ("A", 2), 2: ("A", 4), 3: ("B", 0), 4: ("B", 2), 5: ("B", 4),
           6: ("A", 1), 7: ("A", 3), 8: ("A", 5), 9: ("B", 1), 10: ("B", 3), 11: ("B", 5)}
    at = {v: k for k, v in pos.items()}
    colors = np.array([0] * 6 + [1] * 6)
    ids, tiles = [], []
    for v in range(12):
        body, i = pos[v]
        ring = [v | PREV]                                      # the tile's own start value
        for j in (i - 1, i + 1):
            if (body, j) in at:
                u = at[(body, j)]
                ring.append(u if colors[u] < colors[v] else u | PREV)
        ring += [v] * (32 - len(ring))                         # padding: the tile's first vertex without the flag
        tiles.append([0, v, (1 << 3) | (1 << 9), len(ids)])    # one vertex, one chunk
        ids += ring
    plan = dict(tiles=np.array(tiles, np.uint32), ids=np.array(ids, np.uint32), ctb=np.array([0, 6, 12], np.uint32),
                new2old=np.arange(12, dtype=np.int32), n_active=12, n_colors=2)
    b_verts = np.array([v for v in range(12) if pos[v][0] == "B"])
    contacts = {v: np.random.default_rng(v).choice(b_verts, 6) for v in range(12) if pos[v][0] == "A"}
    kw = dict(favoured={int(v) for v in b_verts}, warp_of=lambda c, j, t: t)
    reads, lead = simulate_contact(plan, colors, 12, 12, np.random.default_rng(0), contacts, lag_bound=True, **kw)
    assert reads == 6 * 6 * 12 and lead == 1                   # as far ahead as the bound allows (sweep k starts when all finished k - 2)
    with pytest.raises(AssertionError, match="history overwritten"):
        simulate_contact(plan, colors, 12, 12, np.random.default_rng(0), contacts, lag_bound=False, **kw)
