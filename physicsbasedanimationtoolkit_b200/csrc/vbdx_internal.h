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
// 32/w vertices, each owned by w adjacent lanes (w = 1,2,4,...,32 chosen from the vertex
// valence so that every lane visits ~tile_iters incident tets).  The incident-tet data of a tile
// is stored as `iters` consecutive *blocks* of 2 KB: block = 4 chunk rows x 32 lanes x 16 B,
// i.e. lane l's 64-byte incidence record is the l-th float4 of each chunk row, so every warp
// load instruction is one fully coalesced 512-byte request.
//
// Incidence record (vertex i, incident tet e), 16 words:
//   word 0..2   internal ids of the three *other* vertices of e (bit 31 set when that vertex has
//               a higher colour than i: it is then read from the previous-iterate buffer, which is
//               what fuses the Chebyshev blend into the sweep)
//   word 3..11  shape-function gradients (rows of GP, fem/ShapeFunctions.h:267-297) of those three
//               vertices, 3 floats each; the gradient of i itself is minus their sum
//   word 12     wg * mu      word 13  wg * lambda      word 14  alpha = 1 + mu/lambda
//   word 15     |grad_i|^2 (precomputed)
// Padding slots have ids = the owning vertex and zero weights, so they contribute exactly 0.
// -----------------------------------------------------------------------------------------
constexpr int kRecordWords      = 16;
constexpr int kBlockFloat4      = 128;  // float4 per block (4 chunk rows x 32 lanes)
constexpr uint32_t kPrevFlag    = 0x80000000u;

struct TileDesc {
    uint32_t blockStart;  // first block of the tile
    uint32_t vbase;       // first internal vertex id
    uint32_t meta;        // log2(w) | iters << 8 | nverts << 24
    uint32_t pad;
};

struct Plan {
    int64_t nV = 0, nActive = 0;
    int32_t nColors = 0;
    std::vector<int32_t> new2old, old2new;   // internal <-> caller vertex ids
    std::vector<TileDesc> tiles;
    std::vector<uint32_t> colorTileBegin;    // nColors + 1
    std::vector<uint32_t> ctaTileRange;      // nColors x (gridBlocks + 1): tiles of colour c for CTA b
    std::vector<uint32_t> ctaBlockBegin;     // nColors x (gridBlocks + 1): first record block of CTA b in colour c
    int64_t nBlocks = 0;
    int64_t nIncidences = 0;                 // over swept vertices
};

// Host planner (plan.cpp).  deg = incident tets per vertex (caller numbering), colors = vertex
// colours, isDbc = Dirichlet mask, X = 3 x nV rest positions (for the Morton order).
void BuildPlan(
    int64_t nV,
    const int32_t* deg,
    const int64_t* colors,
    const uint8_t* isDbc,
    const double* X,
    int tileIters,
    int gridBlocks,
    bool naturalOrder,
    Plan& plan);

// Reference colouring on the host (plan.cpp): graph/Color.h:45-135 over graph/Mesh.h:116-123.
void GreedyColorMesh(
    int64_t nV,
    int64_t nT,
    const int64_t* E,
    int ordering,
    int selection,
    std::vector<int64_t>& colors);

}  // namespace vbdx
