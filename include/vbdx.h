/*
 * vbdx.h -- C ABI of the B200-native Vertex Block Descent integrator.
 *
 * This is the drop-in boundary for the reference's VBD path.  Each entry point names the
 * reference interface it replaces (paths relative to /root/reference/source/pbat unless
 * they start with bindings/).  Conventions at the boundary are the reference's:
 *   - matrices are column-major with one column per vertex / element
 *     (X: 3 x nV, E: 4 x nT, F: 3 x nF;  sim/vbd/Data.h:167-176), i.e. interleaved xyz;
 *   - construction data is double / int64 like pbat::sim::vbd::Data (Aliases.h:17-18);
 *     run-time state crosses as float (GpuScalar, gpu/Aliases.h:19-20) or double;
 *   - the caller owns every host buffer; nothing is retained after a call returns;
 *   - calls on one handle are not thread-safe; every call returns after its work is
 *     complete unless its name ends in _async.
 * Errors: every function returns a vbdx_status; vbdx_last_error() gives the message.  The
 * C++ / Python wrappers rethrow VBDX_INVALID_ARGUMENT as std::invalid_argument / ValueError
 * (the reference throws std::invalid_argument from Data::Construct, sim/vbd/Data.cpp:245-306).
 * There is no CPU fallback: without a CUDA device vbdx_create fails with VBDX_NO_DEVICE.
 */
#ifndef VBDX_H
#define VBDX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VBDX_ABI_VERSION 2

typedef enum vbdx_status {
    VBDX_OK               = 0,
    VBDX_INVALID_ARGUMENT = 1,
    VBDX_NO_DEVICE        = 2,
    VBDX_CUDA_ERROR       = 3,
    VBDX_OUT_OF_MEMORY    = 4,
    VBDX_UNSUPPORTED      = 5
} vbdx_status;

/* sim/vbd/Enums.h:9-15 */
typedef enum vbdx_initialization_strategy {
    VBDX_INIT_POSITION               = 0,
    VBDX_INIT_INERTIA                = 1,
    VBDX_INIT_KINETIC_ENERGY_MINIMUM = 2,
    VBDX_INIT_ADAPTIVE_VBD           = 3,
    VBDX_INIT_ADAPTIVE_PBAT          = 4
} vbdx_initialization_strategy;

/* sim/vbd/Enums.h:21-28.  All six are implemented: the base and Chebyshev solves run inside one persistent launch per
 * step; Anderson, Nesterov, Broyden and TrustRegion wrap one-iteration launches of the same kernel with their device-side
 * window / path kernels. */
typedef enum vbdx_acceleration_strategy {
    VBDX_ACCEL_NONE         = 0,
    VBDX_ACCEL_CHEBYSHEV    = 1,
    VBDX_ACCEL_ANDERSON     = 2,
    VBDX_ACCEL_NESTEROV     = 3,
    VBDX_ACCEL_BROYDEN      = 4,
    VBDX_ACCEL_TRUST_REGION = 5
} vbdx_acceleration_strategy;

/* graph/Enums.h: EGreedyColorOrderingStrategy / EGreedyColorSelectionStrategy */
typedef enum vbdx_color_ordering { VBDX_ORDER_NATURAL = 0, VBDX_ORDER_SMALLEST_DEGREE = 1, VBDX_ORDER_LARGEST_DEGREE = 2 } vbdx_color_ordering;
typedef enum vbdx_color_selection { VBDX_SELECT_LEAST_USED = 0, VBDX_SELECT_FIRST_AVAILABLE = 1 } vbdx_color_selection;

/* How the Chebyshev weight omega_k is computed.  REFERENCE reproduces what
 * sim/vbd/Kernels.h:96-102 evaluates to (omega_k = 1 - rho^2*omega_{k-1} for k >= 2, an operator
 * precedence slip shared by the reference's CPU and GPU integrators); TEXTBOOK is
 * 4/(4 - rho^2*omega_{k-1}). */
typedef enum vbdx_omega_mode { VBDX_OMEGA_REFERENCE = 0, VBDX_OMEGA_TEXTBOOK = 1 } vbdx_omega_mode;

/* hyper-elastic energy of the elastic term.  The reference hard-codes Stable Neo-Hookean
 * (sim/vbd/Integrator.cpp:120, gpu/impl/vbd/Kernels.cuh:178); StVK
 * (physics/SaintVenantKirchhoffEnergy.h) is offered as an extension with the same (mu,lambda). */
typedef enum vbdx_material { VBDX_MATERIAL_STABLE_NEO_HOOKEAN = 0, VBDX_MATERIAL_STVK = 1 } vbdx_material;

/*
 * POD mirror of the fields of pbat::sim::vbd::Data (sim/vbd/Data.h:167-246) that the
 * reference's GPU integrator consumes in its constructor (gpu/impl/vbd/Integrator.cu:22-80,
 * gpu/impl/vbd/ChebyshevIntegrator.cu:16-22).  Not taken from the caller, by design: the
 * vertex->tet adjacency (GVGp/GVGe/GVGilocal), shape-function gradients GP and quadrature
 * weights wg -- they are rebuilt on the device from X and E.
 * Optional pointers may be NULL (defaults = Data::Construct's, sim/vbd/Data.cpp:182-209).
 */
typedef struct vbdx_data_desc {
    uint32_t abi_version;  /* VBDX_ABI_VERSION */
    uint32_t struct_size;  /* sizeof(vbdx_data_desc) */
    int64_t nV, nT;
    const double* X;       /* 3 x nV rest positions (Data::X); also the initial x */
    const int64_t* E;      /* 4 x nT tetrahedra (Data::E) */
    const double* v;       /* 3 x nV initial velocities, default 0 */
    const double* aext;    /* 3 x nV external accelerations, default (0,0,-9.81) */
    const double* m;       /* nV lumped masses; NULL = rho_e*V_e/4 summed on the device */
    const double* rhoe;    /* nT mass densities, default 1e3 (ignored when m is given) */
    const double* lame;    /* 2 x nT (mu_e, lambda_e), default Y=1e6, nu=0.45 */
    const int64_t* dbc;    /* Dirichlet vertices (need not be sorted) */
    int64_t nDbc;
    const int64_t* colors; /* nV vertex colours (Data::colors); NULL = greedy colouring below */
    int32_t ordering;      /* vbdx_color_ordering, reference default LargestDegree */
    int32_t selection;     /* vbdx_color_selection, reference default LeastUsed */
    int32_t strategy;      /* vbdx_initialization_strategy, reference default AdaptivePbat */
    int32_t acceleration;  /* vbdx_acceleration_strategy */
    int32_t omega_mode;    /* vbdx_omega_mode */
    int32_t material;      /* vbdx_material */
    double kD;             /* Rayleigh damping (Data::kD, default 0) */
    double detHZero;       /* Data::detHZero, default 1e-7 */
    double rho;            /* Chebyshev spectral radius estimate, 0 < rho < 1 */
    /* collision mesh (Data::B, V, F) and contact parameters (Data::muC, muF, epsv,
     * mActiveSetUpdateFrequency).  nF == 0 disables contact. */
    const int64_t* B;      /* nV body ids, default all equal */
    const int64_t* V;      /* nCV collision vertices */
    int64_t nCV;
    const int64_t* F;      /* 3 x nF collision triangles */
    int64_t nF;
    double muC, muF, epsv;
    int32_t active_set_update_frequency;
    /* execution */
    int32_t device;        /* CUDA device ordinal, -1 = current device */
    int32_t tile_iters;    /* tuning: target incident tets per lane (0 = default) */
    int32_t flags;         /* VBDX_FLAG_* */
    int32_t kernel_variant;/* tuning: enum vbdx_kernel_variant; 0 = default */
    int32_t ring_slots;    /* tuning: shared-memory ring capacity in 1 KB record blocks (0 = as many as fit) */
    const int64_t* ghosts; /* nGhosts vertices owned by another GPU (domain decomposition): never swept, never touched by
                              the pre-step; their positions are written by the owner.  NULL for a single-GPU problem */
    int64_t nGhosts;
    int32_t consumer_warps;/* tuning: (consumer) warps per CTA of the pipelined / TMA kernels (0 = default) */
    int32_t window_size;   /* Anderson / Broyden acceleration window (Data::mWindowSize, default 5) */
    int32_t n_colors;      /* colours of the WHOLE problem when this handle simulates a part of it (domain decomposition: every
                              rank must sweep the same number of colours); 0 = largest colour of this handle's vertices + 1 */
    int32_t nesterov_start;/* Data::mNesterovAccelerationStart (sim/vbd/Data.h:239; default 3) */
    double nesterov_L;     /* Data::mNesterovLipschitzConstant (sim/vbd/Data.h:238; default 1) */
    double tr_eta, tr_tau; /* trust-region acceptance ratio and radius growth factor (sim/vbd/Data.h:241-242; defaults 0.2, 2) */
    int32_t tr_curved;     /* Data::bCurved: curved (default) or linear accelerated path (sim/vbd/Data.h:243) */
    int32_t reserved0;
} vbdx_data_desc;

/* Which persistent step kernel runs the sweeps.  Both compute the same arithmetic in the same order. */
typedef enum vbdx_kernel_variant {
    VBDX_KERNEL_DEFAULT = 0,
    VBDX_KERNEL_DIRECT  = 1, /* every warp loads its records straight from global memory */
    VBDX_KERNEL_TMA     = 2, /* warp-specialised: a producer warp streams records into a shared-memory ring with
                                bulk asynchronous copies (TMA) across the colour barriers */
    VBDX_KERNEL_PIPELINED = 3, /* every warp prefetches its next tile's static data with cp.async while it computes
                                 the current one (the default when the per-warp buffers fit in shared memory) */
    VBDX_KERNEL_CLUSTER = 4   /* small meshes: the whole problem is swept by ONE thread-block cluster (8 CTAs x 16 warps)
                                 and colours are separated by the hardware cluster barrier instead of a grid barrier
                                 through L2 (by default the shape of the launches that keep colour barriers on such meshes:
                                 partial launches, VBDX_DATAFLOW=0; whole steps run the barrier-free kernel) */
} vbdx_kernel_variant;

#define VBDX_FLAG_ADAPTIVE_VBD_GPU_HISTORY 1 /* AdaptiveVbd uses the stored v(t-1) like the reference's GPU
                                                path (gpu/impl/vbd/Integrator.cu:333) instead of the CPU path's
                                                vt == v (sim/vbd/Integrator.cpp:32,61-68) */
#define VBDX_FLAG_NATURAL_VERTEX_ORDER 2     /* keep vertices in colour-major natural order (no Morton) */

typedef struct vbdx_integrator vbdx_integrator;

/* Fill a descriptor with the reference's defaults (sim/vbd/Data.h:206-246). */
void vbdx_data_desc_init(vbdx_data_desc* desc);

/* pbat::gpu::vbd::Integrator::Integrator(Data const&)   gpu/vbd/Integrator.h:45, gpu/vbd/Integrator.cu:15-33
 * pbat::sim::vbd::Integrator::Integrator(Data)          sim/vbd/Integrator.h:22 */
vbdx_status vbdx_create(const vbdx_data_desc* desc, vbdx_integrator** out);
/* Independent scenes (no coupling, no communication) stepped together: SURVEY.md 8e "independent scene batches",
 * BASELINE configs[4].  descs[0..n) describe one scene each and must agree in the solver settings; the handle then
 * behaves like an integrator over the concatenation of the scenes (scene s owns vertices
 * [vertex_offsets[s], vertex_offsets[s+1]) of every get/set call).  A scene in a batch evolves bit-identically to the
 * same scene stepped alone.  No reference counterpart (the reference steps one Data per Integrator). */
vbdx_status vbdx_create_batch(const vbdx_data_desc* descs, int32_t n, vbdx_integrator** out);
/* n_scenes and/or the n_scenes + 1 vertex offsets of a batch handle (either pointer may be NULL) */
vbdx_status vbdx_batch_offsets(vbdx_integrator* h, int32_t* n_scenes, int64_t* vertex_offsets);
/* ~Integrator   gpu/vbd/Integrator.h:62 */
vbdx_status vbdx_destroy(vbdx_integrator* h);

/* Integrator::Step(dt, iterations, substeps)   gpu/vbd/Integrator.h:72, sim/vbd/Integrator.h:29
 * One persistent launch per call (with contact: per substep, behind the active-set kernels, which are replayed as CUDA
 * graphs).  The colours of a sweep follow each other without a grid barrier (every position carries the number of its write; a
 * tile waits for exactly its 1-ring; contact terms read other bodies' vertices from a history of their last four writes), with
 * results bit-identical to the barrier sweep; environment: VBDX_DATAFLOW=0 selects colour barriers, VBDX_DATAFLOW_TIMEOUT_S (2) bounds a dependency
 * wait (exceeded = VBDX_CUDA_ERROR with a diagnostic, and the handle falls back to barriers). */
vbdx_status vbdx_step(vbdx_integrator* h, double dt, int32_t iterations, int32_t substeps);
/* same, returns once the work is enqueued on the handle's stream */
vbdx_status vbdx_step_async(vbdx_integrator* h, double dt, int32_t iterations, int32_t substeps);
vbdx_status vbdx_synchronize(vbdx_integrator* h);

/* Integrator::SetPositions / SetVelocities / SetExternalAcceleration   gpu/vbd/Integrator.h:95-105
 * (float = the GPU wrapper's GpuMatrixX; double = sim::vbd::Integrator's data.x / data.v,
 *  bindings/pypbat/sim/vbd/Integrator.cpp:60-71).  3 x nV column-major, caller's vertex order. */
vbdx_status vbdx_set_positions_f32(vbdx_integrator* h, const float* x, int64_t nV);
vbdx_status vbdx_set_positions_f64(vbdx_integrator* h, const double* x, int64_t nV);
vbdx_status vbdx_set_velocities_f32(vbdx_integrator* h, const float* v, int64_t nV);
vbdx_status vbdx_set_velocities_f64(vbdx_integrator* h, const double* v, int64_t nV);
vbdx_status vbdx_set_external_acceleration_f32(vbdx_integrator* h, const float* a, int64_t nV);
vbdx_status vbdx_set_external_acceleration_f64(vbdx_integrator* h, const double* a, int64_t nV);
/* Integrator::GetPositions / GetVelocities   gpu/vbd/Integrator.h:139-144 */
vbdx_status vbdx_get_positions_f32(vbdx_integrator* h, float* x, int64_t nV);
vbdx_status vbdx_get_positions_f64(vbdx_integrator* h, double* x, int64_t nV);
vbdx_status vbdx_get_velocities_f32(vbdx_integrator* h, float* v, int64_t nV);
vbdx_status vbdx_get_velocities_f64(vbdx_integrator* h, double* v, int64_t nV);

/* The same setters/getters with an explicit element type and storage order, so that a caller holding a
 * row-major 3 x nV array (numpy's default for the arrays bindings/pypbat hands out) needs no host-side
 * transposition: VBDX_LAYOUT_COLUMNS = column-major 3 x nV (Eigen; xyz of a vertex adjacent),
 * VBDX_LAYOUT_ROWS = row-major 3 x nV (all x, then all y, then all z).  Host pointers may be pageable or
 * pinned; with pinned memory (vbdx_host_alloc) the copy is a single DMA transfer. */
typedef enum vbdx_field {
    VBDX_FIELD_POSITIONS = 0, VBDX_FIELD_VELOCITIES = 1, VBDX_FIELD_EXTERNAL_ACCELERATION = 2,
    VBDX_FIELD_INERTIAL_TARGET = 3,   /* read-only: Data::xtilde of the current substep */
    VBDX_FIELD_PREVIOUS_POSITIONS = 4 /* read-only: Data::xt */
} vbdx_field;
typedef enum vbdx_dtype { VBDX_F32 = 0, VBDX_F64 = 1 } vbdx_dtype;
typedef enum vbdx_layout { VBDX_LAYOUT_COLUMNS = 0, VBDX_LAYOUT_ROWS = 1 } vbdx_layout;
vbdx_status vbdx_set_vertex_field(vbdx_integrator* h, int32_t field, int32_t dtype, int32_t layout, const void* src, int64_t nV);
vbdx_status vbdx_get_vertex_field(vbdx_integrator* h, int32_t field, int32_t dtype, int32_t layout, void* dst, int64_t nV);
/* The same, enqueued on the handle's stream without waiting (extension, like vbdx_step_async): the host array should be
 * page-locked (vbdx_host_alloc; a pageable one makes the copy synchronous) and must not be touched -- read after a get,
 * written after a set -- before vbdx_synchronize returns.  set_async, step_async, get_async, synchronize is one host
 * round trip per step instead of three. */
vbdx_status vbdx_set_vertex_field_async(vbdx_integrator* h, int32_t field, int32_t dtype, int32_t layout, const void* src, int64_t nV);
vbdx_status vbdx_get_vertex_field_async(vbdx_integrator* h, int32_t field, int32_t dtype, int32_t layout, void* dst, int64_t nV);
/* Page-locked host memory for the arrays above (cudaHostAlloc / cudaFreeHost). */
vbdx_status vbdx_host_alloc(void** out, int64_t bytes);
vbdx_status vbdx_host_free(void* p);

/* A slice of ONE substep of length sdt, for callers that look at every iterate (Integrator::TraceNextStep /
 * ExportTrace, sim/vbd/Integrator.cpp:47-52,202-235; gpu TracedStep, gpu/impl/vbd/Integrator.cu:105-148,284-301):
 * [pre-step: xt, xtilde, initial guess] -> iterations k_begin..k_end-1 of a solve of total_iterations ->
 * [post-step: velocity update].  Base and Chebyshev solves; blocking. */
#define VBDX_PARTIAL_PRE_STEP 1
#define VBDX_PARTIAL_POST_STEP 2
vbdx_status vbdx_step_partial(vbdx_integrator* h, double sdt, int32_t k_begin, int32_t k_end, int32_t total_iterations, int32_t flags);
/* Integrator::ObjectiveFunction / ObjectiveFunctionGradient (sim/vbd/Integrator.h:45-58, Integrator.cpp:138-200):
 * f = 1/2 |xk - xtilde|_M^2 + dt^2 sum_e wg_e psi_e(xk), evaluated on the device in double precision.
 * xk, xtilde: 3 x nV column-major, caller's vertex order; f and grad (3 nV) may each be NULL. */
vbdx_status vbdx_objective(vbdx_integrator* h, const double* xk, const double* xtilde, double dt, double* f, double* grad);

/* Integrator::SetNumericalZeroForHessianDeterminant   gpu/vbd/Integrator.h:111 */
vbdx_status vbdx_set_detH_zero(vbdx_integrator* h, double zero);
/* Integrator::SetRayleighDampingCoefficient           gpu/vbd/Integrator.h:116 */
vbdx_status vbdx_set_rayleigh_damping(vbdx_integrator* h, double kD);
/* Integrator::SetInitializationStrategy               gpu/vbd/Integrator.h:121 */
vbdx_status vbdx_set_initialization_strategy(vbdx_integrator* h, int32_t strategy);
/* Integrator::SetBlockSize                            gpu/vbd/Integrator.h:126
 * accepted for compatibility; the sweep is warp-tiled, so this is only a tuning hint */
vbdx_status vbdx_set_block_size(vbdx_integrator* h, int32_t block_size);
/* Extension (BASELINE north star: "fused 3x3 Newton solve with line-search guard"); no reference counterpart.
 * 0 (default) = the reference's semantics: the full Newton step is always taken (sim/vbd/Kernels.h:329-339).
 * 1 = guarded step: a step that is not a finite descent direction of the vertex' local objective is replaced by a
 *     scaled steepest-descent step, and with the St. Venant-Kirchhoff energy (no damping / contact) the step length
 *     is chosen by Armijo backtracking over t in {1, 1/2, 1/4} on the true local objective (none passes: the vertex
 *     stays).  With the Stable Neo-Hookean energy the local objective is exactly quadratic and the Newton step is
 *     its minimiser, so the guard never alters a step (tested: bit-identical results). */
vbdx_status vbdx_set_line_search_guard(vbdx_integrator* h, int32_t enabled);
/* Integrator::SetSceneBoundingBox                     gpu/vbd/Integrator.h:132-134 */
vbdx_status vbdx_set_scene_bounding_box(vbdx_integrator* h, const float min3[3], const float max3[3]);

/* ---- stand-alone LBVH and vertex-triangle detector ------------------------------------------
 * pbat::gpu::geometry::Bvh (gpu/geometry/Bvh.h; bindings/pypbat/gpu/geometry/Bvh.cpp:19-109).  Boxes are 3 x n
 * column-major float arrays; node arrays use the reference's numbering (internal nodes 0..n-2, leaves n-1..2n-2 in
 * Morton order). */
typedef struct vbdx_bvh vbdx_bvh;
vbdx_status vbdx_bvh_create(int64_t max_boxes, vbdx_bvh** out);                                   /* Bvh(max_boxes, .) */
vbdx_status vbdx_bvh_destroy(vbdx_bvh* h);
vbdx_status vbdx_bvh_build(vbdx_bvh* h, int64_t n, const float* lo, const float* hi, const float wmin[3], const float wmax[3]); /* Bvh::Build */
/* child 2 x (n-1), parent 2n-1, rightmost 2 x (n-1), inds n, codes n, node boxes 3 x (2n-1), visits n-1; any may be NULL */
vbdx_status vbdx_bvh_get(vbdx_bvh* h, int32_t* child, int32_t* parent, int32_t* rightmost, int32_t* inds, uint32_t* codes, float* node_lo,
                         float* node_hi, int32_t* visits);
/* Bvh::DetectOverlaps: self-overlaps (bi < bj) of the boxes given to build(); with set != NULL only pairs with
 * set[bi] != set[bj].  pairs is 2 x max_overlaps (column per pair); *n_found may exceed max_overlaps (the rest is dropped) */
vbdx_status vbdx_bvh_detect_overlaps(vbdx_bvh* h, const int32_t* set, int64_t max_overlaps, int32_t* pairs, int64_t* n_found);
/* Bvh::PointTriangleNearestNeighbors: nearest triangle of nQ points; the tree was built over the nF triangles' boxes */
vbdx_status vbdx_bvh_nearest_triangles(vbdx_bvh* h, int64_t nQ, const float* X, int64_t nP, const float* V, int64_t nF, const int32_t* F, int32_t* out);

/* pbat::gpu::contact::VertexTriangleMixedCcdDcd (gpu/contact/VertexTriangleMixedCcdDcd.h;
 * bindings/pypbat/gpu/contact/VertexTriangleMixedCcdDcd.cpp:18-82).  Positions are 3 x nV column-major float. */
typedef struct vbdx_contact vbdx_contact;
vbdx_status vbdx_contact_create(int64_t nV, const int64_t* B, const int64_t* V, int64_t nCV, const int64_t* F, int64_t nF, vbdx_contact** out);
vbdx_status vbdx_contact_destroy(vbdx_contact* h);
vbdx_status vbdx_contact_initialize_active_set(vbdx_contact* h, const float* xt, const float* xtp1, const float wmin[3], const float wmax[3]);
vbdx_status vbdx_contact_update_active_set(vbdx_contact* h, const float* x);
vbdx_status vbdx_contact_finalize_active_set(vbdx_contact* h, const float* x);
vbdx_status vbdx_contact_set_eps(vbdx_contact* h, float eps);
/* active mask (nCV), nearest triangles (nCV x 8, -1 terminated), compacted active vertices (first *n_active of nCV) */
vbdx_status vbdx_contact_get(vbdx_contact* h, int32_t* active_mask, int32_t* nn, int32_t* av, int64_t* n_active);

/* Host-only debug access to the sweep plan (tiles, ring lists with their previous-iterate flags and padding, colour
 * ranges, internal numbering) of a mesh: what the CPU test-suite model-checks the barrier-free sweep protocol against.
 * is_constrained: 0 = swept, 1 = Dirichlet, 2 = ghost.  vbdx_debug_plan_get(what): 0 sizes {nTiles, nRingIds, nColors,
 * nActive, ghostBegin} (int64 x 5), 1 tiles (uint32 x 4 each: blockStart, vbase, meta, ringStart), 2 ring ids (uint32),
 * 3 colour tile begins (uint32 x (nColors + 1)), 4 new2old (int32 x nV). */
typedef struct vbdx_plan vbdx_plan;
vbdx_status vbdx_debug_plan_create(int64_t nV, int64_t nT, const int64_t* E, const int64_t* colors, const uint8_t* is_constrained,
                                   const double* X, int32_t tile_iters, vbdx_plan** out);
vbdx_status vbdx_debug_plan_get(vbdx_plan* h, int32_t what, void* out);
vbdx_status vbdx_debug_plan_destroy(vbdx_plan* h);

/* Use a caller-provided cudaStream_t for all subsequent work (NULL = the handle's own). */
vbdx_status vbdx_set_stream(vbdx_integrator* h, void* cuda_stream);

/* Introspection used by tests, bench.py and the roofline accounting. */
typedef struct vbdx_info {
    int64_t nV, nT, nActiveVertices; /* nActiveVertices = vertices that are swept (non-Dirichlet) */
    int64_t nIncidences;             /* sum over swept vertices of incident tets */
    int64_t nRecordSlots;            /* incidence record slots incl. padding (x 32 B = streamed bytes/sweep) */
    int32_t nColors, nTiles;
    int32_t gridBlocks, blockThreads; /* persistent launch shape */
    int32_t device, smCount;
    int64_t deviceBytes;             /* device memory held by the handle */
    int64_t kernelLaunches;          /* CUDA kernels launched by this handle since creation */
    double  lastStepMs;              /* device time of the last vbdx_step (CUDA events on its stream) */
    int64_t nRingEntries;            /* sum over tiles of the distinct vertices they stage (own vertices + 1-rings, unpadded) */
    int64_t nGhosts;                 /* vertices owned by another GPU (domain decomposition) */
    int64_t nonFiniteVertices;       /* sentinel: owned vertices whose position was NaN/Inf at the end of the last step (0 = healthy) */
} vbdx_info;
vbdx_status vbdx_get_info(vbdx_integrator* h, vbdx_info* out);

/* Debug / parity access to what the device built (caller's vertex and element numbering):
 * the vertex->tet CSR (GVGp nV+1, GVGe and GVGilocal 4 nT; sim/vbd/Data.h:196-204), shape function
 * gradients GP (4 x 3nT as Data::GP), quadrature weights wg (nT), lumped masses m (nV), and
 * colours (nV).  Any pointer may be NULL. */
vbdx_status vbdx_get_adjacency(vbdx_integrator* h, int64_t* GVGp, int64_t* GVGe, int64_t* GVGilocal);
vbdx_status vbdx_get_element_data(vbdx_integrator* h, double* GP, double* wg, double* m);
vbdx_status vbdx_get_colors(vbdx_integrator* h, int64_t* colors);

/* ---- multi-GPU domain decomposition (new: the reference is single-GPU) --------------------
 * One handle per GPU/process.  Each handle simulates the vertices it owns plus a ghost layer (desc->ghosts).  When
 * a boundary vertex is swept, its owner stores the new position directly into the ghost slots of its peers over
 * NVLink (peer-to-peer stores inside the persistent step kernel).  Every such store carries the number of the
 * write ("tag") in the fourth component of the same 16 bytes, and a reader waits until the ghost it gathered
 * carries the tag of the write it needs -- so the halo exchange needs no fence, flag or collective on the critical
 * path; the GPUs only bound how far they may drift apart (two colour phases).
 *   1. vbdx_get_internal_ids: caller-order vertex -> device-internal slot (peers need the slots of their ghosts)
 *   2. vbdx_dist_ipc_handles: 128 opaque bytes to all-gather between the processes (CUDA IPC handles)
 *   3. vbdx_dist_connect: open the peers' buffers and install the send lists:
 *        send_local[k]  caller-order id of an owned vertex,
 *        send_peer[k]   rank that holds it as a ghost,
 *        send_remote[k] that rank's internal slot of the ghost;
 *        peer_nverts[r] / peer_nghosts[r] = rank r's local vertex count / how many of them are ghosts;
 *        recv_mask bit r = rank r owns some of this rank's ghosts.  A GPU synchronises only with the ranks it
 *        sends to or receives from.
 * All ranks must then call vbdx_step with identical arguments. */
vbdx_status vbdx_get_internal_ids(vbdx_integrator* h, int64_t* old2new);
vbdx_status vbdx_dist_ipc_handles(vbdx_integrator* h, void* out128);
vbdx_status vbdx_dist_connect(vbdx_integrator* h, int32_t rank, int32_t world, const void* all_handles, const int64_t* peer_nverts,
                              const int64_t* peer_nghosts, int64_t nSend, const int64_t* send_local, const int64_t* send_peer, const int64_t* send_remote,
                              uint32_t recv_mask);

/* Diagnostics of the halo exchange since the last reset: out4 = {ghost values that had not arrived when a tile needed
 * them, ns spent polling them (summed over lanes), colour barriers (CTA 0) that had to wait for a neighbour's epoch,
 * ns spent there}. */
vbdx_status vbdx_dist_stats(vbdx_integrator* h, uint32_t out4[4], int32_t reset);

/* Contact state after the last step, per collision vertex in the order of desc->V
 * (gpu/impl/contact/VertexTriangleMixedCcdDcd.cuh: active, nn): active[nCV] (0/1), nn[8 * nCV] nearest triangles
 * (indices into desc->F, -1 terminated), *nActive.  Any pointer may be NULL. */
vbdx_status vbdx_get_contact_state(vbdx_integrator* h, int32_t* active, int32_t* nn, int64_t* nActive);

/* Stand-alone LBVH build over n boxes (lo, hi: 3 x n column-major) inside the world box [wmin, wmax], as
 * pbat::gpu::impl::geometry::Bvh::Build does (gpu/impl/geometry/Bvh.cu:130-137).  Outputs use the reference's
 * conventions: child 2 x (n-1) and rightmost 2 x (n-1) stored as [left row | right row], parent 2n-1, nodes
 * 0..n-2 internal (root 0) and n-1..2n-2 leaves in sorted order, inds = leaf -> box.  Any output may be NULL. */
vbdx_status vbdx_debug_bvh_build(int64_t n, const float* lo, const float* hi, const float wmin[3], const float wmax[3], int32_t* child,
                                 int32_t* parent, int32_t* rightmost, int32_t* inds, uint32_t* codes, float* nodeLo, float* nodeHi);

/* Diagnostics: with out == NULL, arm phase tracing of sweep `iteration` for the following steps (direct kernel
 * variant); with out != NULL, read back nColors x gridBlocks x 8 %globaltimer stamps (phase start, warp 0 done,
 * CTA done, barrier released; warp 0's first tile: descriptor loaded, 1-rings staged, tets accumulated, solved) and disarm. */
vbdx_status vbdx_debug_trace(vbdx_integrator* h, int32_t iteration, unsigned long long* out, int64_t capacity);

/* Host-only helpers (no device needed): the reference's greedy colouring of the mesh primal
 * graph, graph/Color.h:45-135 on graph/Mesh.h:116-123, as Data::Construct calls it
 * (sim/vbd/Data.cpp:228-231). */
vbdx_status vbdx_greedy_color(int64_t nV, int64_t nT, const int64_t* E, int32_t ordering, int32_t selection, int64_t* colors_out);

const char* vbdx_last_error(void);
/* ---- XPBD over the same contact pipeline -------------------------------------------------
 * Replaces pbat::gpu::xpbd::Integrator (gpu/xpbd/Integrator.h:27-170; impl gpu/impl/xpbd/Integrator.cu:88-448), constructed from
 * what pbat::sim::xpbd::Data holds after Construct() (sim/xpbd/Data.h:19-104, sim/xpbd/Data.cpp:102-160).  Arrays follow the
 * reference's conventions: X 3 x nV and T 4 x nT column-major, Pptr / Padj (and SGptr, SGadj, Cptr, Cadj for clustered
 * partitions) compressed sparse lists of constraint (= element) ids.  One persistent cooperative launch runs a whole
 * substep (a whole Step without a collision mesh); partitions are separated by grid barriers inside the kernel. */
typedef enum vbdx_xpbd_constraint { VBDX_XPBD_STABLE_NEO_HOOKEAN = 0, VBDX_XPBD_COLLISION = 1 } vbdx_xpbd_constraint; /* sim/xpbd/Enums.h:9-11 */
typedef struct vbdx_xpbd_desc {
    uint32_t abi_version, struct_size;
    int64_t nV, nT;
    const double* X;        /* 3 x nV particle positions (Data::x) */
    const int64_t* T;       /* 4 x nT tetrahedra */
    const double* v;        /* 3 x nV or NULL (zero) */
    const double* aext;     /* 3 x nV or NULL (gravity -9.81 along z, sim/xpbd/Data.cpp:108-112) */
    const double* minv;     /* nV inverse masses or NULL (1e-3, sim/xpbd/Data.cpp:113-116) */
    const double* lame;     /* 2 x nT or NULL (Y = 1e6, nu = 0.45, sim/xpbd/Data.cpp:127-133) */
    const int64_t* dbc;     /* Dirichlet vertices: minv = 0, v = a = 0 (sim/xpbd/Data.cpp:122-125) */
    int64_t nDbc;
    const int64_t* Pptr;    /* nPartitions + 1 */
    const int64_t* Padj;
    int32_t nPartitions, nClusterPartitions;
    const int64_t *SGptr, *SGadj, *Cptr, *Cadj; /* clustered partitions (Data::WithClusterPartitions) or NULL */
    const double* alphaSNH; /* 2 x nT compliances or NULL (1 / (lame * volume), sim/xpbd/Data.cpp:145-148) */
    const double* betaSNH;  /* 2 x nT damping or NULL (0) */
    const int64_t* BV;      /* nV body ids or NULL (one body) */
    const int64_t* V;       /* nCV collision vertices */
    int64_t nCV;
    const int64_t* F;       /* 3 x nF collision triangles */
    int64_t nF;
    const double* muV;      /* nCV collision penalties or NULL (1) */
    const double* alphaC;   /* nCV contact compliances / damping or NULL (0); per collision vertex (the reference indexes them by slot */
    const double* betaC;    /*   in its active list, gpu/impl/xpbd/Integrator.cu:372-414: identical for the uniform values it is used with) */
    double muS, muD;        /* static / dynamic friction (sim/xpbd/Data.h:85-86) */
    int32_t active_set_update_frequency;
    int32_t device;
} vbdx_xpbd_desc;
typedef struct vbdx_xpbd vbdx_xpbd;
void vbdx_xpbd_desc_init(vbdx_xpbd_desc* d);
vbdx_status vbdx_xpbd_create(const vbdx_xpbd_desc* desc, vbdx_xpbd** out);
vbdx_status vbdx_xpbd_destroy(vbdx_xpbd* h);
vbdx_status vbdx_xpbd_step(vbdx_xpbd* h, double dt, int32_t iterations, int32_t substeps);      /* gpu/xpbd/Integrator.h:66 */
vbdx_status vbdx_xpbd_set_positions(vbdx_xpbd* h, const double* x, int64_t nV);                   /* :77 */
vbdx_status vbdx_xpbd_set_velocities(vbdx_xpbd* h, const double* v, int64_t nV);                  /* :82 */
vbdx_status vbdx_xpbd_set_external_acceleration(vbdx_xpbd* h, const double* a, int64_t nV);       /* :87 */
vbdx_status vbdx_xpbd_get_positions(vbdx_xpbd* h, double* x, int64_t nV);                         /* :71 */
vbdx_status vbdx_xpbd_get_velocities(vbdx_xpbd* h, double* v, int64_t nV);
vbdx_status vbdx_xpbd_set_compliance(vbdx_xpbd* h, int32_t constraint, const double* alpha, int64_t n); /* :119 */
vbdx_status vbdx_xpbd_set_friction_coefficients(vbdx_xpbd* h, double muS, double muD);            /* :126 */
vbdx_status vbdx_xpbd_set_scene_bounding_box(vbdx_xpbd* h, const float min3[3], const float max3[3]); /* :133 */
/* out8 = nV, nT, partitions, grid blocks, kernel launches, device bytes, last step in ns, nCV */
vbdx_status vbdx_xpbd_get_info(vbdx_xpbd* h, int64_t* out8);
vbdx_status vbdx_xpbd_get_contact_state(vbdx_xpbd* h, int32_t* active, int32_t* nn, int64_t* nActive);
/* graph/Color.h:45-135 on a graph in compressed sparse format (bindings/pypbat/graph/Color.cpp:28-60 greedy_color); host only */
vbdx_status vbdx_graph_greedy_color(int64_t n, const int64_t* ptr, const int64_t* adj, int32_t ordering, int32_t selection, int64_t* colors_out);
/* The same colouring computed on the device, for the selection that parallelises (FirstAvailable = 1: the smallest colour
 * missing among the neighbours that come earlier in the visiting order depends on those neighbours only, so vertices are
 * coloured in rounds as soon as theirs are -- the sequential result, vertex for vertex; rounds_out: how many rounds).
 * LeastUsed (0), the reference's default, picks by a global running count and is inherently sequential: VBDX_UNSUPPORTED.
 * E: 4 x nT column-major like vbdx_greedy_color (sim/vbd/Data.cpp:228-231 -> graph/Color.h:45-135). */
vbdx_status vbdx_greedy_color_device(int64_t nV, int64_t nT, const int64_t* E, int32_t ordering, int32_t selection, int32_t device,
                                     int64_t* colors_out, int32_t* rounds_out);

/* Test hooks (GPU): the sweep's vertex-triangle contact term (csrc/contact.cuh, restating sim/vbd/Kernels.h:223-302) and its
 * area-scaled penalties (gpu/impl/vbd/Kernels.cuh:80-114) on caller-supplied inputs.  in28 = per pair xtv(3) xv(3) xtf(3 x 3,
 * one triangle vertex after the other) xf(3 x 3) dt k muF epsv;  out13 = 0, g(3), H(3 x 3).  fc = 8 triangle ids per vertex. */
vbdx_status vbdx_debug_contact_pairs(int32_t n, const float* in28, float* out13);
vbdx_status vbdx_debug_contact_penalties(int32_t nVerts, int32_t nTris, const int32_t* fc, const float* XVA, const float* FA, float muC,
                                         int32_t* nContacts, float* penalty);
int32_t vbdx_abi_version(void);
/* number of CUDA devices visible (0 when there is no driver/GPU) */
int32_t vbdx_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VBDX_H */
