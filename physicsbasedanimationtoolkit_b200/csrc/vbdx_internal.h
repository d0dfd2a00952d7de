// Internal declarations shared by the host planner, the device setup kernels, the persistent
// step kernel and the C-ABI layer.  Nothing here is part of the public boundary (include/vbdx.h).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace vbdx {

// -----------------------------------------------------------------------------------------
// Data layout in HBM (see DESIGN.md "Data layout")
//
// Vertices are renumbered ("internal ids"): swept vertices first, colour-major, inside a colour
// grouped into warp tiles; Dirichlet vertices last.  A *tile* is the unit of work of one warp:
// up to 32/w vertices, each owned by w adjacent lanes (w = 1,2,4,...,32 chosen from the vertex
// valence so that every lane visits ~tile_iters incident tets).
//
// Per tile two static streams exist:
//  * the *ring list*: the tile's own vertices followed by the distinct vertices of their 1-rings,
//    as internal ids (bit 31 set when the entry is read from the previous-iterate buffer: own
//    vertices and neighbours of a higher colour -- which is what fuses the Chebyshev blend into
//    the sweep).  The warp gathers these positions ONCE per tile into shared memory; every
//    incident tet then addresses its three other vertices by 10-bit local indices.  Neighbours
//    whose colour is the one swept immediately before the tile's colour come last ("late"
//    chunks): everything else is stable during the preceding colour and can be gathered before
//    the barrier.
//  * the *incidence records*, `iters` consecutive blocks of 1 KB: block = 2 chunk rows x 32
//    lanes x 16 B, i.e. lane l's 32-byte record is the l-th float4 of each chunk row, so every
//    warp load instruction is one fully coalesced 512-byte request.
//
// Incidence record (vertex i, incident tet e with other vertices a, b, c), 8 words -- everything
// the closed-form Stable Neo-Hookean vertex block needs (DESIGN.md "Math"):
//   word 0   local ring indices of a, b, c, 10 bits each
//   word 1-3 wg*mu*(u_a+u_b+u_c), wg*mu*u_b, wg*mu*u_c   with u_n = grad N_n . grad N_i
//   word 4   beta  = wg*lambda*detG^2        (detG = det of the rows grad N_a, grad N_b, grad N_c)
//   word 5   gamma = wg*lambda*detG*alpha    (alpha = 1 + mu/lambda)
//   word 6   wg*mu*|grad N_i|^2              word 7  unused (0)
// Padding slots are all-zero and therefore contribute exactly 0.
//
// St. Venant-Kirchhoff needs the deformation gradient itself, i.e. the three shape-function gradients of the other
// vertices: its incidence record is TWO consecutive blocks (same lane), so that every copy path stays as it is:
//   block 2t   : word 0 local ring indices | grad N_a (3) || grad N_b (3) | wg*mu
//   block 2t+1 : grad N_c (3) | wg*lambda  || unused (0)
// TileIters then counts blocks (2 per incident tet).
// -----------------------------------------------------------------------------------------
constexpr int kRecordWords      = 8;
constexpr int kBlockFloat4      = 64;   // float4 per block (2 chunk rows x 32 lanes)
constexpr int kBlockBytes       = kBlockFloat4 * 16;
constexpr uint32_t kPrevFlag    = 0x80000000u;
constexpr int kMaxRingPerTile   = 1024; // 10-bit local indices

struct TileDesc {
    uint32_t blockStart;  // first record block of the tile
    uint32_t vbase;       // first internal vertex id
    uint32_t meta;        // log2(w) [0:3) | nverts [3:9) | ring chunks [9:15) | early ring chunks [15:21) | iters [21:31) | reads ghosts [31]
    uint32_t ringStart;   // first entry of the tile's ring list (multiple of 32)
};

#if defined(__CUDACC__)
#define VBDX_HD __host__ __device__
#else
#define VBDX_HD
#endif
VBDX_HD constexpr uint32_t TileLog2W(uint32_t meta) { return meta & 7u; }
VBDX_HD constexpr uint32_t TileVerts(uint32_t meta) { return (meta >> 3) & 63u; }
VBDX_HD constexpr uint32_t TileChunks(uint32_t meta) { return (meta >> 9) & 63u; }       // all ring chunks of 32 entries
VBDX_HD constexpr uint32_t TileEarlyChunks(uint32_t meta) { return (meta >> 15) & 63u; } // chunks that do not depend on the previous colour
VBDX_HD constexpr uint32_t TileIters(uint32_t meta) { return (meta >> 21) & 1023u; }
VBDX_HD constexpr bool TileReadsGhosts(uint32_t meta) { return (meta >> 31) != 0u; }  // its ring list holds vertices owned by another GPU
VBDX_HD constexpr uint32_t TileMeta(uint32_t lw, uint32_t nverts, uint32_t chunks, uint32_t early, uint32_t iters)
{
    return lw | (nverts << 3) | (chunks << 9) | (early << 15) | (iters << 21);
}
constexpr uint32_t kMaxTileIters = 1023;

struct Plan {
    int64_t nV = 0, nActive = 0;
    int64_t ghostBegin = 0;                  // internal ids >= ghostBegin are ghosts of another GPU's vertices
    int32_t nColors = 0;
    std::vector<int32_t> new2old, old2new;   // internal <-> caller vertex ids
    std::vector<TileDesc> tiles;
    std::vector<uint32_t> colorTileBegin;    // nColors + 1
    std::vector<uint32_t> ctaTileRange;      // nColors x (gridBlocks + 1): tiles of colour c for CTA b
    std::vector<uint32_t> ctaBlockBegin;     // nColors x (gridBlocks + 1): first record block of CTA b in colour c
    std::vector<uint32_t> ringIds;           // ring lists of all tiles (internal ids | kPrevFlag)
    std::vector<uint32_t> recIdx;            // per record slot: packed local ring indices of the three other vertices
    int32_t maxRingPerTile = 0;              // longest (padded) tile ring list: sizes the per-warp staging
    int64_t nBlocks = 0;
    int64_t nIncidences = 0;                 // over swept vertices
    int64_t nRingEntries = 0;                // sum of 1-ring sizes over swept vertices (unpadded)
};

// Host planner (plan.cpp).  E / vtPtr / vtAdj = connectivity (caller numbering), colors = vertex
// colours, isDbc = Dirichlet mask, X = 3 x nV rest positions (for the Morton order).
void BuildPlan(
    int64_t nV,
    const int32_t* E,        // 4 x nT
    const uint32_t* vtPtr,   // vertex -> tet CSR built on the device (nV + 1)
    const uint32_t* vtAdj,   // entries 4*e + ilocal
    const int64_t* colors,
    const uint8_t* isDbc,    // 1 = Dirichlet, 2 = ghost (both are never swept)
    const double* X,
    int tileIters,
    bool naturalOrder,
    int blocksPerIncidence,  // record blocks per incident tet: 1 (Stable Neo-Hookean), 2 (St. Venant-Kirchhoff)
    int minColors,           // sweep at least this many colours (domain decomposition: the colour count of the whole mesh)
    Plan& plan);

// Second planning step, once the persistent grid size is known: which tiles / record blocks of each
// colour belong to which CTA.
void PartitionTiles(Plan& plan, int gridBlocks);

// Reference colouring on the host (plan.cpp): graph/Color.h:45-135 over graph/Mesh.h:116-123.
void GreedyColorMesh(
    int64_t nV,
    int64_t nT,
    const int64_t* E,
    int ordering,
    int selection,
    std::vector<int64_t>& colors);

// ... and on any graph in compressed sparse format (bindings/pypbat/graph/Color.cpp:28-60 greedy_color)
void GreedyColorGraph(int64_t n, const int64_t* ptr, const int32_t* adj, int ordering, int selection, std::vector<int64_t>& colors);

}  // namespace vbdx
