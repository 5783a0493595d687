// warp-per-game kernel, float state
#include "hk_generic_launch.inl"
namespace hk {
int launch_generic_f32(const StepParams& p, bool obs, int dev, cudaStream_t stream) {
    return obs ? dispatch_generic<float, true>(p, dev, stream) : dispatch_generic<float, false>(p, dev, stream);
}
}  // namespace hk
