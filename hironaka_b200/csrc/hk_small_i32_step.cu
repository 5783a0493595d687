// thread-per-game kernel, int32_t state, instantiations WITHOUT observation features
#include "hk_small_launch.inl"
namespace hk {
int launch_small_i32_noobs(const StepParams& p, int dev, cudaStream_t stream) { return dispatch_small<int32_t, false>(p, dev, stream); }
}  // namespace hk
