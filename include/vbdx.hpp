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

  private:
    static void Check(vbdx_status s)
    {
        if (s == VBDX_OK)
            return;
        std::string const what = vbdx_last_error();
        if (s == VBDX_INVALID_ARGUMENT)
            throw std::invalid_argument(what);
        throw std::runtime_error(what);
    }
    vbdx_integrator* mImpl{nullptr};
    int64_t mNV{0};
};

}  // namespace vbd
}  // namespace gpu
}  // namespace pbat_b200

#endif  // VBDX_HPP
