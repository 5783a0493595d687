// thread-per-game kernel, float state, instantiations WITH observation features
#include "hk_small_launch.inl"
namespace hk {
int launch_small_f32_obs(const StepParams& p, int dev, cudaStream_t stream) { return dispatch_small<float, true>(p, dev, stream); }
}  // namespace hk
