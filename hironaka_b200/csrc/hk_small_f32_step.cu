// thread-per-game kernel, float state, instantiations WITHOUT observation features
#include "hk_small_launch.inl"
namespace hk {
int launch_small_f32_noobs(const StepParams& p, int dev, cudaStream_t stream) { return dispatch_small<float, false>(p, dev, stream); }
}  // namespace hk
