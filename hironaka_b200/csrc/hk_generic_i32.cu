// warp-per-game kernel, int32_t state
#include "hk_generic_launch.inl"
namespace hk {
int launch_generic_i32(const StepParams& p, bool obs, int dev, cudaStream_t stream) {
    return obs ? dispatch_generic<int32_t, true>(p, dev, stream) : dispatch_generic<int32_t, false>(p, dev, stream);
}
}  // namespace hk
