// vbdx.hpp -- header-only C++ class over the C ABI (include/vbdx.h) with the method names of
// pbat::gpu::vbd::Integrator (source/pbat/gpu/vbd/Integrator.h:33-148) and
// pbat::sim::vbd::Integrator (source/pbat/sim/vbd/Integrator.h:15-63), so that code written against
// the reference compiles against this class by changing the namespace and passing raw 3 x nV
// column-major arrays where the reference takes Eigen matrices (Eigen is not a dependency here).
// Error behaviour follows the reference: std::invalid_argument for ill-formed input, and
// std::runtime_error for device errors.
#ifndef VBDX_HPP
#define VBDX_HPP

#include "vbdx.h"

#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace pbat_b200 {
namespace gpu {
namespace vbd {

enum class EInitializationStrategy { Position, Inertia, KineticEnergyMinimum, AdaptiveVbd, AdaptivePbat };

class Integrator
{
  public:
    /// Construct from the POD mirror of pbat::sim::vbd::Data (gpu/vbd/Integrator.h:45)
    explicit Integrator(vbdx_data_desc const& data) { Check(vbdx_create(&data, &mImpl)); mNV = data.nV; }
    Integrator(Integrator const&)            = delete;  // gpu/vbd/Integrator.h:48-49
    Integrator& operator=(Integrator const&) = delete;
    Integrator(Integrator&& o) noexcept : mImpl(std::exchange(o.mImpl, nullptr)), mNV(o.mNV) {}
    Integrator& operator=(Integrator&& o) noexcept
    {
        if (this != &o)
        {
            vbdx_destroy(mImpl);
            mImpl = std::exchange(o.mImpl, nullptr);
            mNV   = o.mNV;
        }
        return *this;
    }
    ~Integrator() { vbdx_destroy(mImpl); }

    /// gpu/vbd/Integrator.h:72, sim/vbd/Integrator.h:29
    void Step(float dt, int iterations, int substeps = 1) { Check(vbdx_step(mImpl, dt, iterations, substeps)); }
    /// gpu/vbd/Integrator.h:95-105; x, v, aext are 3 x nV column-major (xyz interleaved)
    void SetPositions(float const* x) { Check(vbdx_set_positions_f32(mImpl, x, mNV)); }
    void SetVelocities(float const* v) { Check(vbdx_set_velocities_f32(mImpl, v, mNV)); }
    void SetExternalAcceleration(float const* a) { Check(vbdx_set_external_acceleration_f32(mImpl, a, mNV)); }
    void SetPositions(double const* x) { Check(vbdx_set_positions_f64(mImpl, x, mNV)); }
    void SetVelocities(double const* v) { Check(vbdx_set_velocities_f64(mImpl, v, mNV)); }
    /// gpu/vbd/Integrator.h:111-134
    void SetNumericalZeroForHessianDeterminant(float zero) { Check(vbdx_set_detH_zero(mImpl, zero)); }
    void SetRayleighDampingCoefficient(float kD) { Check(vbdx_set_rayleigh_damping(mImpl, kD)); }
    void SetInitializationStrategy(EInitializationStrategy s) { Check(vbdx_set_initialization_strategy(mImpl, static_cast<int>(s))); }
    /// extension: guarded Newton step (include/vbdx.h vbdx_set_line_search_guard); off = the reference's behaviour
    void SetLineSearchGuard(bool enabled) { Check(vbdx_set_line_search_guard(mImpl, enabled ? 1 : 0)); }
    void SetBlockSize(int blockSize) { Check(vbdx_set_block_size(mImpl, blockSize)); }
    void SetSceneBoundingBox(float const min3[3], float const max3[3]) { Check(vbdx_set_scene_bounding_box(mImpl, min3, max3)); }
    /// gpu/vbd/Integrator.h:139-144: 3 x nV, returned by value
    std::vector<float> GetPositions() const
    {
        std::vector<float> x(3 * mNV);
        Check(vbdx_get_positions_f32(mImpl, x.data(), mNV));
        return x;
    }
    std::vector<float> GetVelocities() const
    {
        std::vector<float> v(3 * mNV);
        Check(vbdx_get_velocities_f32(mImpl, v.data(), mNV));
        return v;
    }
    vbdx_integrator* Handle() const { return mImpl; }

    static void Check(vbdx_status s)
    {
        if (s == VBDX_OK)
            return;
        std::string const what = vbdx_last_error();
        if (s == VBDX_INVALID_ARGUMENT)
            throw std::invalid_argument(what);
        throw std::runtime_error(what);
    }

  protected:
    explicit Integrator(vbdx_integrator* adopted) : mImpl(adopted) {}
    void SetVertexCount(int64_t nV) { mNV = nV; }

  private:
    vbdx_integrator* mImpl{nullptr};
    int64_t mNV{0};
};

/// Independent scenes behind one handle (vbdx_create_batch): an Integrator over the concatenation of the scenes.
class BatchIntegrator : public Integrator
{
  public:
    explicit BatchIntegrator(std::vector<vbdx_data_desc> const& scenes) : Integrator(Create(scenes))
    {
        int32_t n = 0;
        vbdx_batch_offsets(Handle(), &n, nullptr);
        mOffsets.resize(static_cast<std::size_t>(n) + 1);
        vbdx_batch_offsets(Handle(), nullptr, mOffsets.data());
        SetVertexCount(mOffsets.back());
    }
    /// first vertex of every scene, then the total
    std::vector<int64_t> const& Offsets() const { return mOffsets; }

  private:
    static vbdx_integrator* Create(std::vector<vbdx_data_desc> const& scenes)
    {
        vbdx_integrator* h = nullptr;
        Check(vbdx_create_batch(scenes.data(), static_cast<int32_t>(scenes.size()), &h));
        return h;
    }
    std::vector<int64_t> mOffsets;
};

}  // namespace vbd

namespace geometry {

/// pbat::gpu::geometry::Bvh (gpu/geometry/Bvh.h:33-148) over raw 3 x n column-major float boxes
class Bvh
{
  public:
    Bvh(int64_t maxBoxes, int64_t maxOverlaps) : mMaxOverlaps(maxOverlaps) { vbd::Integrator::Check(vbdx_bvh_create(maxBoxes, &mImpl)); }
    Bvh(Bvh const&)            = delete;
    Bvh& operator=(Bvh const&) = delete;
    ~Bvh() { vbdx_bvh_destroy(mImpl); }
    void Build(int64_t n, float const* lo, float const* hi, float const wmin[3], float const wmax[3])
    {
        vbd::Integrator::Check(vbdx_bvh_build(mImpl, n, lo, hi, wmin, wmax));
        mN = n;
    }
    /// 2 x #overlaps (column per pair, bi < bj); set == nullptr: all pairs, else only pairs from different sets
    std::vector<int32_t> DetectOverlaps(int32_t const* set = nullptr)
    {
        std::vector<int32_t> pairs(2 * static_cast<std::size_t>(mMaxOverlaps > 0 ? mMaxOverlaps : 1));
        int64_t found = 0;
        vbd::Integrator::Check(vbdx_bvh_detect_overlaps(mImpl, set, mMaxOverlaps, pairs.data(), &found));
        pairs.resize(2 * static_cast<std::size_t>(found < mMaxOverlaps ? found : mMaxOverlaps));
        return pairs;
    }
    std::vector<int32_t> PointTriangleNearestNeighbors(int64_t nQ, float const* X, int64_t nP, float const* V, int32_t const* F)
    {
        std::vector<int32_t> nn(static_cast<std::size_t>(nQ));
        vbd::Integrator::Check(vbdx_bvh_nearest_triangles(mImpl, nQ, X, nP, V, mN, F, nn.data()));
        return nn;
    }
    vbdx_bvh* Handle() const { return mImpl; }

  private:
    vbdx_bvh* mImpl{nullptr};
    int64_t mMaxOverlaps{0}, mN{0};
};

}  // namespace geometry

namespace contact {

/// pbat::gpu::contact::VertexTriangleMixedCcdDcd (gpu/contact/VertexTriangleMixedCcdDcd.h) over raw arrays
class VertexTriangleMixedCcdDcd
{
  public:
    VertexTriangleMixedCcdDcd(int64_t nV, int64_t const* B, int64_t const* V, int64_t nCV, int64_t const* F, int64_t nF) : mNCV(nCV)
    {
        vbd::Integrator::Check(vbdx_contact_create(nV, B, V, nCV, F, nF, &mImpl));
    }
    VertexTriangleMixedCcdDcd(VertexTriangleMixedCcdDcd const&)            = delete;
    VertexTriangleMixedCcdDcd& operator=(VertexTriangleMixedCcdDcd const&) = delete;
    ~VertexTriangleMixedCcdDcd() { vbdx_contact_destroy(mImpl); }
    void InitializeActiveSet(float const* xt, float const* xtp1, float const wmin[3], float const wmax[3])
    {
        vbd::Integrator::Check(vbdx_contact_initialize_active_set(mImpl, xt, xtp1, wmin, wmax));
    }
    void UpdateActiveSet(float const* x) { vbd::Integrator::Check(vbdx_contact_update_active_set(mImpl, x)); }
    void FinalizeActiveSet(float const* x) { vbd::Integrator::Check(vbdx_contact_finalize_active_set(mImpl, x)); }
    void SetNearestNeighbourFloatingPointTolerance(float eps) { vbd::Integrator::Check(vbdx_contact_set_eps(mImpl, eps)); }
    /// indices (into V) of the active vertices
    std::vector<int32_t> ActiveVertices() const
    {
        std::vector<int32_t> av(static_cast<std::size_t>(mNCV));
        int64_t n = 0;
        vbd::Integrator::Check(vbdx_contact_get(mImpl, nullptr, nullptr, av.data(), &n));
        av.resize(static_cast<std::size_t>(n));
        return av;
    }

  private:
    vbdx_contact* mImpl{nullptr};
    int64_t mNCV{0};
};

}  // namespace contact

namespace xpbd {

/// Constraint kinds of pbat::sim::xpbd::EConstraint (sim/xpbd/Enums.h)
enum class EConstraint { StableNeoHookean = 0, Collision = 1 };

/// pbat::gpu::xpbd::Integrator (gpu/xpbd/Integrator.h:33-140) over raw 3 x nV column-major double arrays
class Integrator
{
  public:
    explicit Integrator(vbdx_xpbd_desc const& data) : mNV(data.nV) { vbd::Integrator::Check(vbdx_xpbd_create(&data, &mImpl)); }
    Integrator(Integrator const&)            = delete;
    Integrator& operator=(Integrator const&) = delete;
    ~Integrator() { vbdx_xpbd_destroy(mImpl); }
    void Step(double dt, int iterations, int substeps) { vbd::Integrator::Check(vbdx_xpbd_step(mImpl, dt, iterations, substeps)); }
    void SetPositions(double const* x) { vbd::Integrator::Check(vbdx_xpbd_set_positions(mImpl, x, mNV)); }
    void SetVelocities(double const* v) { vbd::Integrator::Check(vbdx_xpbd_set_velocities(mImpl, v, mNV)); }
    void SetExternalAcceleration(double const* a) { vbd::Integrator::Check(vbdx_xpbd_set_external_acceleration(mImpl, a, mNV)); }
    void SetCompliance(double const* alpha, int64_t n, EConstraint c)
    {
        vbd::Integrator::Check(vbdx_xpbd_set_compliance(mImpl, static_cast<int32_t>(c), alpha, n));
    }
    void SetFrictionCoefficients(double muS, double muD) { vbd::Integrator::Check(vbdx_xpbd_set_friction_coefficients(mImpl, muS, muD)); }
    void SetSceneBoundingBox(float const min3[3], float const max3[3])
    {
        vbd::Integrator::Check(vbdx_xpbd_set_scene_bounding_box(mImpl, min3, max3));
    }
    std::vector<double> GetPositions() const
    {
        std::vector<double> x(static_cast<std::size_t>(3 * mNV));
        vbd::Integrator::Check(vbdx_xpbd_get_positions(mImpl, x.data(), mNV));
        return x;
    }
    std::vector<double> GetVelocities() const
    {
        std::vector<double> v(static_cast<std::size_t>(3 * mNV));
        vbd::Integrator::Check(vbdx_xpbd_get_velocities(mImpl, v.data(), mNV));
        return v;
    }

  private:
    vbdx_xpbd* mImpl{nullptr};
    int64_t mNV{0};
};

}  // namespace xpbd
}  // namespace gpu

namespace graph {

/// pbat::graph::GreedyColor (graph/Color.h:45-135) on a graph in compressed sparse format; host only
inline std::vector<int64_t> GreedyColor(std::vector<int64_t> const& ptr, std::vector<int64_t> const& adj, int ordering = 2, int selection = 0)
{
    std::vector<int64_t> colors(ptr.empty() ? 0 : ptr.size() - 1);
    gpu::vbd::Integrator::Check(
        vbdx_graph_greedy_color(static_cast<int64_t>(colors.size()), ptr.data(), adj.data(), ordering, selection, colors.data()));
    return colors;
}

}  // namespace graph
}  // namespace pbat_b200

#endif  // VBDX_HPP
