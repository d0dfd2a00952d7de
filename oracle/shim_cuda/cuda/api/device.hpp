// Stand-in for the third-party cuda-api-wrappers headers (<cuda/api/...>, absent here; the reference fetches them through
// vcpkg) that gpu/impl/vbd/Kernels.cuh includes for its host-side launcher `Invoke`.  Only what that template needs to
// PARSE is declared; contact_ref.cu launches the reference's kernels with plain <<< >>> and never calls it.
#pragma once
namespace cuda {
namespace memory { namespace shared { using size_t = unsigned; } }
struct launch_configuration_t {};
struct launch_config_builder {
    launch_config_builder& block_size(int) { return *this; }
    launch_config_builder& dynamic_shared_memory_size(unsigned) { return *this; }
    launch_config_builder& grid_size(int) { return *this; }
    launch_configuration_t build() { return {}; }
};
namespace device {
struct device_t {
    template <class K, class... A>
    void launch(K, launch_configuration_t, A&&...) {}
};
namespace current { inline device_t get() { return {}; } }
}  // namespace device
}  // namespace cuda
