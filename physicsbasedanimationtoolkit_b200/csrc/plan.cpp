// Host-side planning for the B200 VBD integrator: vertex colouring (when the caller supplies
// none) and the colour-major / warp-tile vertex order the sweep kernel streams through.
// Pure C++ (no CUDA) so that it can be unit-tested on a CPU-only machine through the C-ABI.
#include "vbdx_internal.h"

#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <numeric>
#include <stdexcept>

namespace vbdx {

namespace {

constexpr uint32_t kSelfMarker = 0xfffffffeu;
constexpr uint32_t kPadMarker  = 0xffffffffu;

inline uint32_t ExpandBits10(uint32_t v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

}  // namespace

// Greedy colouring of the mesh primal graph with the reference's semantics
// (graph/Color.h:45-135; visiting order = stable sort by vertex degree in the primal graph
// G*G^T, self loop included, graph/Mesh.h:116-123; colour choice = usable colour with the fewest
// vertices, ties to the lowest colour index; a new colour only when all are blocked).
void GreedyColorMesh(
    int64_t nV,
    int64_t nT,
    const int64_t* E,
    int ordering,
    int selection,
    std::vector<int64_t>& colors)
{
    // vertex -> tet lists by counting sort
    std::vector<int64_t> vtPtr(nV + 1, 0);
    for (int64_t k = 0; k < 4 * nT; ++k)
        ++vtPtr[E[k] + 1];
    for (int64_t i = 0; i < nV; ++i)
        vtPtr[i + 1] += vtPtr[i];
    std::vector<int32_t> vt(static_cast<size_t>(4 * nT));
    {
        std::vector<int64_t> cursor(vtPtr.begin(), vtPtr.end() - 1);
        for (int64_t e = 0; e < nT; ++e)
            for (int a = 0; a < 4; ++a)
                vt[cursor[E[4 * e + a]]++] = static_cast<int32_t>(e);
    }
    // 1-ring (self included) via a visit stamp; two passes: sizes, then entries
    std::vector<int64_t> ringPtr(nV + 1, 0);
    std::vector<int64_t> stamp(nV, -1);
    for (int64_t u = 0; u < nV; ++u)
    {
        int64_t cnt = 0;
        stamp[u]    = u;
        ++cnt;
        for (int64_t k = vtPtr[u]; k < vtPtr[u + 1]; ++k)
            for (int a = 0; a < 4; ++a)
            {
                int64_t const w = E[4 * static_cast<int64_t>(vt[k]) + a];
                if (stamp[w] != u)
                {
                    stamp[w] = u;
                    ++cnt;
                }
            }
        ringPtr[u + 1] = ringPtr[u] + cnt;
    }
    std::vector<int32_t> ring(static_cast<size_t>(ringPtr[nV]));
    std::fill(stamp.begin(), stamp.end(), int64_t(-1));
    for (int64_t u = 0; u < nV; ++u)
    {
        int64_t pos = ringPtr[u];
        stamp[u]    = u;
        ring[pos++] = static_cast<int32_t>(u);
        for (int64_t k = vtPtr[u]; k < vtPtr[u + 1]; ++k)
            for (int a = 0; a < 4; ++a)
            {
                int64_t const w = E[4 * static_cast<int64_t>(vt[k]) + a];
                if (stamp[w] != u)
                {
                    stamp[w]    = u;
                    ring[pos++] = static_cast<int32_t>(w);
                }
            }
    }
    GreedyColorGraph(nV, ringPtr.data(), ring.data(), ordering, selection, colors);
}

// graph/Color.h:45-135 on a graph in compressed sparse format (ptr, adj): vertices are visited in natural order or in
// (stable) order of their degree ptr[u+1] - ptr[u]; a vertex takes, among the colours none of its neighbours has, the
// one with the fewest vertices (LeastUsed; ties to the lowest index) or the lowest one (FirstAvailable); a new colour is
// opened only when all are blocked.
void GreedyColorGraph(int64_t n, const int64_t* ptr, const int32_t* adj, int ordering, int selection, std::vector<int64_t>& colors)
{
    std::vector<int64_t> order(n);
    if (ordering == 0 /* Natural */)
        std::iota(order.begin(), order.end(), int64_t(0));
    else
    {
        int64_t maxDeg = 0;
        for (int64_t u = 0; u < n; ++u)
            maxDeg = std::max(maxDeg, ptr[u + 1] - ptr[u]);
        std::vector<int64_t> bucket(maxDeg + 2, 0);
        bool const largestFirst = (ordering == 2);
        auto slot = [&](int64_t u) {
            int64_t const d = ptr[u + 1] - ptr[u];
            return largestFirst ? (maxDeg - d) : d;
        };
        for (int64_t u = 0; u < n; ++u)
            ++bucket[slot(u) + 1];
        for (int64_t d = 0; d <= maxDeg; ++d)
            bucket[d + 1] += bucket[d];
        for (int64_t u = 0; u < n; ++u)
            order[bucket[slot(u)]++] = u;
    }
    colors.assign(n, -1);
    std::vector<int64_t> used;     // vertices per colour
    std::vector<uint8_t> blocked;  // per colour, for the vertex being coloured
    for (int64_t u : order)
    {
        std::fill(blocked.begin(), blocked.end(), uint8_t(0));
        size_t nBlocked = 0;
        for (int64_t k = ptr[u]; k < ptr[u + 1]; ++k)
        {
            int64_t const c = colors[adj[k]];
            if (c >= 0 && !blocked[c])
            {
                blocked[c] = 1;
                ++nBlocked;
            }
        }
        if (nBlocked == used.size())
        {
            colors[u] = static_cast<int64_t>(used.size());
            used.push_back(1);
            blocked.push_back(0);
            continue;
        }
        int64_t pick = -1;
        for (int64_t c = 0; c < static_cast<int64_t>(used.size()); ++c)
        {
            if (blocked[c])
                continue;
            if (pick < 0 || (selection == 0 /* LeastUsed */ && used[c] < used[pick]))
                pick = c;
        }
        colors[u] = pick;
        ++used[pick];
    }
}

void BuildPlan(
    int64_t nV,
    const int32_t* E,
    const uint32_t* vtPtr,
    const uint32_t* vtAdj,
    const int64_t* colors,
    const uint8_t* isDbc,
    const double* X,
    int tileIters,
    bool naturalOrder,
    int blocksPerIncidence,
    int minColors,
    Plan& plan)
{
    plan      = Plan{};
    blocksPerIncidence = std::max(1, blocksPerIncidence);
    plan.nV   = nV;
    tileIters = std::max(1, tileIters);
    // bounding box for the Morton order
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = 0; i < nV; ++i)
        for (int d = 0; d < 3; ++d)
        {
            lo[d] = std::min(lo[d], X[3 * i + d]);
            hi[d] = std::max(hi[d], X[3 * i + d]);
        }
    double ext = 0;
    for (int d = 0; d < 3; ++d)
        ext = std::max(ext, hi[d] - lo[d]);
    if (!(ext > 0))
        ext = 1;
    // 1-ring (distinct other vertices of the incident tets) of every vertex, caller numbering
    std::vector<uint32_t> ringPtr(nV + 1, 0);
    std::vector<int32_t> ring;
    {
        std::vector<int64_t> stamp(nV, -1);
        ring.reserve(static_cast<size_t>(nV) * 14);
        for (int64_t u = 0; u < nV; ++u)
        {
            if (!isDbc[u])
            {
                stamp[u] = u;
                for (uint32_t k = vtPtr[u]; k < vtPtr[u + 1]; ++k)
                {
                    int64_t const e = vtAdj[k] >> 2;
                    for (int a = 0; a < 4; ++a)
                    {
                        int32_t const w = E[4 * e + a];
                        if (stamp[w] != u)
                        {
                            stamp[w] = u;
                            ring.push_back(w);
                        }
                    }
                }
            }
            ringPtr[u + 1] = static_cast<uint32_t>(ring.size());
        }
    }
    const char* boEnv       = std::getenv("VBDX_BOUNDARY_ORDER");
    int const boundaryOrder = boEnv ? std::atoi(boEnv) : 0;
    struct Item {
        uint64_t key;
        int32_t v;
        uint8_t lw;
        uint16_t iters;
    };
    std::vector<Item> items;
    items.reserve(nV);
    int64_t nColors = std::max(0, minColors);
    for (int64_t i = 0; i < nV; ++i)
    {
        nColors = std::max<int64_t>(nColors, colors[i] + 1);
        if (isDbc[i])
            continue;
        int const d = static_cast<int>(vtPtr[i + 1] - vtPtr[i]);
        int lw      = 0;
        while (lw < 5 && (int64_t(1) << lw) * tileIters < d)
            ++lw;
        int const w     = 1 << lw;
        int const iters = std::max(1, (d + w - 1) / w);
        uint32_t morton = 0;
        if (!naturalOrder)
        {
            uint32_t q[3];
            for (int k = 0; k < 3; ++k)
                q[k] = static_cast<uint32_t>(
                    std::min(1023.0, std::max(0.0, (X[3 * i + k] - lo[k]) / ext * 1024.0)));
            morton = (ExpandBits10(q[0]) << 2) | (ExpandBits10(q[1]) << 1) | ExpandBits10(q[2]);
        }
        uint64_t const itKey = 4095u - static_cast<uint32_t>(std::min(iters, 4095));
        // domain decomposition tuning knob (VBDX_BOUNDARY_ORDER = 1 / 2): vertices next to another GPU's vertices first /
        // last within their colour.  Default 0 (Morton order): both groupings measured slower.
        uint64_t group = 0;
        if (boundaryOrder != 0)
        {
            bool boundary = false;
            for (uint32_t r = ringPtr[i]; r < ringPtr[i + 1]; ++r)
                boundary |= isDbc[ring[r]] == 2;
            group = (boundaryOrder == 1) ? (boundary ? 0u : 1u) : (boundary ? 1u : 0u);
        }
        uint64_t const key = (static_cast<uint64_t>(colors[i]) << 46) | (group << 45) |
                             (static_cast<uint64_t>(5 - lw) << 42) | (itKey << 30) | morton;
        items.push_back({key, static_cast<int32_t>(i), static_cast<uint8_t>(lw),
                         static_cast<uint16_t>(std::min(iters, 65535))});
        plan.nIncidences += d;
    }
    std::stable_sort(items.begin(), items.end(), [](Item const& a, Item const& b) { return a.key < b.key; });
    plan.nActive = static_cast<int64_t>(items.size());
    plan.nColors = static_cast<int32_t>(nColors);
    plan.new2old.resize(nV);
    plan.old2new.resize(nV);
    plan.colorTileBegin.assign(nColors + 1, 0);
    // internal numbering: swept vertices in sorted order (tiles take consecutive runs), Dirichlet vertices last
    for (int64_t k = 0; k < plan.nActive; ++k)
    {
        plan.new2old[k]              = items[k].v;
        plan.old2new[items[k].v]     = static_cast<int32_t>(k);
    }
    {
        int64_t tail = plan.nActive;
        for (int pass = 1; pass <= 2; ++pass)  // Dirichlet vertices, then ghosts
        {
            if (pass == 2)
                plan.ghostBegin = tail;
            for (int64_t i = 0; i < nV; ++i)
                if (isDbc[i] == pass)
                {
                    plan.new2old[tail] = static_cast<int32_t>(i);
                    plan.old2new[i]    = static_cast<int32_t>(tail);
                    ++tail;
                }
        }
    }
    // tiles
    std::vector<int64_t> seenInTile(nV, -1);  // neighbour -> tile that already lists it
    std::vector<uint32_t> localIndex(nV, 0);
    std::vector<int32_t> early, late;
    int64_t pos = 0, block = 0;
    int64_t curColor = 0;
    while (pos < plan.nActive)
    {
        int64_t const c = colors[items[pos].v];
        while (curColor < c)
            plan.colorTileBegin[++curColor] = static_cast<uint32_t>(plan.tiles.size());
        int64_t const prevColor = (c + nColors - 1) % nColors;  // the colour swept right before c
        int64_t const tileId    = static_cast<int64_t>(plan.tiles.size());
        int const lw = items[pos].lw;
        int const G  = 32 >> lw;
        int n = 0, iters = 0;
        early.clear();
        late.clear();
        while (n < G && pos + n < plan.nActive && colors[items[pos + n].v] == c &&
               items[pos + n].lw == lw)
        {
            int32_t const v = items[pos + n].v;
            // distinct neighbours this vertex adds to the tile's list
            size_t const e0 = early.size(), l0 = late.size();
            for (uint32_t r = ringPtr[v]; r < ringPtr[v + 1]; ++r)
            {
                int32_t const j = ring[r];
                if (seenInTile[j] == tileId)
                    continue;
                seenInTile[j] = tileId;
                (nColors > 1 && colors[j] == prevColor && isDbc[j] != 1 ? late : early).push_back(j);  // ghosts (2) do change
            }
            size_t const padded = ((n + 1 + early.size() + 31) / 32 + (late.size() + 31) / 32) * 32;
            if (n > 0 && padded > static_cast<size_t>(kMaxRingPerTile))
            {
                // keep the tile's list addressable with 10-bit local indices: undo and close the tile
                for (size_t k = e0; k < early.size(); ++k)
                    seenInTile[early[k]] = -1;
                for (size_t k = l0; k < late.size(); ++k)
                    seenInTile[late[k]] = -1;
                early.resize(e0);
                late.resize(l0);
                break;
            }
            iters = std::max<int>(iters, items[pos + n].iters);
            ++n;
        }
        uint32_t const earlyChunks = static_cast<uint32_t>((n + early.size() + 31) / 32);
        uint32_t const lateChunks  = static_cast<uint32_t>((late.size() + 31) / 32);
        if ((earlyChunks + lateChunks) * 32 > static_cast<uint32_t>(kMaxRingPerTile))
            throw std::length_error("a vertex has more than ~1000 distinct neighbours");
        if (iters * blocksPerIncidence > static_cast<int>(kMaxTileIters))
            throw std::length_error("a vertex has too many incident tetrahedra for one warp tile");
        uint32_t const ringStart = static_cast<uint32_t>(plan.ringIds.size());
        TileDesc t;
        t.blockStart = static_cast<uint32_t>(block);
        t.vbase      = static_cast<uint32_t>(pos);
        t.meta       = TileMeta(static_cast<uint32_t>(lw), static_cast<uint32_t>(n), earlyChunks + lateChunks, earlyChunks,
                          static_cast<uint32_t>(iters * blocksPerIncidence));
        t.ringStart = ringStart;
        for (int32_t j : early)
            if (isDbc[j] == 2)
                t.meta |= 0x80000000u;
        for (int32_t j : late)
            if (isDbc[j] == 2)
                t.meta |= 0x80000000u;
        plan.tiles.push_back(t);
        // neighbours in ascending internal id: lanes of one gather instruction then touch few cache lines
        auto byNewId = [&](int32_t a, int32_t b) { return plan.old2new[a] < plan.old2new[b]; };
        std::sort(early.begin(), early.end(), byNewId);
        std::sort(late.begin(), late.end(), byNewId);
        // list layout: [own vertices | early neighbours | pad] [late neighbours | pad]; caller ids for now
        plan.ringIds.resize(ringStart + static_cast<size_t>(earlyChunks + lateChunks) * 32, kPadMarker);
        for (int k = 0; k < n; ++k)
            plan.ringIds[ringStart + k] = kSelfMarker;
        for (size_t k = 0; k < early.size(); ++k)
        {
            plan.ringIds[ringStart + n + k] = static_cast<uint32_t>(early[k]);
            localIndex[early[k]]            = static_cast<uint32_t>(n + k);
        }
        for (size_t k = 0; k < late.size(); ++k)
        {
            plan.ringIds[ringStart + earlyChunks * 32 + k] = static_cast<uint32_t>(late[k]);
            localIndex[late[k]]                             = static_cast<uint32_t>(earlyChunks * 32 + k);
        }
        plan.nRingEntries += static_cast<int64_t>(n + early.size() + late.size());
        plan.maxRingPerTile = std::max<int32_t>(plan.maxRingPerTile, static_cast<int32_t>((earlyChunks + lateChunks) * 32));
        // packed local indices of every record slot of the tile (same slot enumeration as FillRecords)
        plan.recIdx.resize(static_cast<size_t>(block + iters * blocksPerIncidence) * 32, 0u);
        uint32_t const w = 1u << lw;
        for (int t2 = 0; t2 < iters; ++t2)
            for (uint32_t lane = 0; lane < 32; ++lane)
            {
                uint32_t const grp = lane >> lw, sub = lane & (w - 1u);
                if (grp >= static_cast<uint32_t>(n))
                    continue;
                int32_t const v  = items[pos + grp].v;
                uint32_t const k = static_cast<uint32_t>(t2) * w + sub;
                if (k >= vtPtr[v + 1] - vtPtr[v])
                    continue;
                uint32_t const packed = vtAdj[vtPtr[v] + k];
                int64_t const e       = packed >> 2;
                uint32_t const il     = packed & 3u;
                uint32_t idx = 0;
                int m        = 0;
                for (uint32_t a = 0; a < 4; ++a)
                    if (a != il)
                        idx |= localIndex[E[4 * e + a]] << (10 * m++);
                plan.recIdx[static_cast<size_t>(block + t2 * blocksPerIncidence) * 32 + lane] = idx;
            }
        pos += n;
        block += iters * blocksPerIncidence;
    }
    while (curColor < nColors)
        plan.colorTileBegin[++curColor] = static_cast<uint32_t>(plan.tiles.size());
    plan.nBlocks = block;
    // ring lists: caller ids -> internal ids; flag = read from the previous-iterate buffer (own vertices, and
    // neighbours with a higher colour than the tile's)
    for (size_t t = 0; t < plan.tiles.size(); ++t)
    {
        TileDesc const& td = plan.tiles[t];
        uint32_t const nv  = TileVerts(td.meta);
        uint32_t const end = td.ringStart + TileChunks(td.meta) * 32u;
        int64_t const c    = colors[plan.new2old[td.vbase]];
        for (uint32_t r = td.ringStart; r < end; ++r)
        {
            uint32_t const raw = plan.ringIds[r];
            if (raw == kSelfMarker)
                plan.ringIds[r] = (td.vbase + (r - td.ringStart)) | kPrevFlag;
            else if (raw == kPadMarker)
                plan.ringIds[r] = td.vbase;  // harmless load
            else
            {
                uint32_t id = static_cast<uint32_t>(plan.old2new[raw]);
                if (colors[raw] > c)
                    id |= kPrevFlag;
                plan.ringIds[r] = id;
            }
        }
        (void)nv;
    }
}

// Per colour, split the tile range between the CTAs of the persistent grid by equal record-block counts.
void PartitionTiles(Plan& plan, int gridBlocks)
{
    int64_t const nColors = plan.nColors;
    gridBlocks = std::max(1, gridBlocks);
    plan.ctaTileRange.assign(static_cast<size_t>(nColors) * (gridBlocks + 1), 0);
    plan.ctaBlockBegin.assign(static_cast<size_t>(nColors) * (gridBlocks + 1), 0);
    for (int64_t c = 0; c < nColors; ++c)
    {
        uint32_t const tb = plan.colorTileBegin[c], te = plan.colorTileBegin[c + 1];
        uint32_t* range   = &plan.ctaTileRange[c * (gridBlocks + 1)];
        int64_t const b0  = (tb < te) ? plan.tiles[tb].blockStart : 0;
        int64_t const b1  = (tb < te) ? (te < plan.tiles.size() ? plan.tiles[te].blockStart : plan.nBlocks) : 0;
        int64_t const total = b1 - b0;
        uint32_t t = tb;
        for (int b = 0; b < gridBlocks; ++b)
        {
            range[b] = t;
            int64_t const limit = b0 + (total * (b + 1)) / gridBlocks;
            while (t < te && static_cast<int64_t>(plan.tiles[t].blockStart) < limit)
                ++t;
        }
        range[gridBlocks] = te;
        uint32_t* blk = &plan.ctaBlockBegin[c * (gridBlocks + 1)];
        for (int b = 0; b <= gridBlocks; ++b)
            blk[b] = static_cast<uint32_t>(range[b] < te ? plan.tiles[range[b]].blockStart : b1);
    }
}

}  // namespace vbdx
