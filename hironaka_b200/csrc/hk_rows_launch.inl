// hk_rows_launch.inl — launch code of the small-games kernel of large padded shapes, included by hk_rows_*.cu
#include "hk_launch.cuh"
#include "hk_rows.cuh"

namespace hk {
namespace {

template <typename T, int D>
int launch_rows_d(const StepParams& p, int dev, cudaStream_t stream) {
    constexpr int WARPS = 8, STAGES = 2;
    using L = RowsLayout<D, WARPS, STAGES>;
    static KernelFacts facts;
    auto kernel = hk_rows_kernel<T, D, WARPS, STAGES>;
    cudaError_t err = cudaSuccess;
    const int per_sm = kernel_ctas_per_sm(kernel, facts, dev, WARPS * 32, L::SMEM_BYTES, &err);
    if (err != cudaSuccess) return (int)err;
    long long ctas = (p.B + WARPS * 32 - 1) / (WARPS * 32);
    const long long cap = (long long)device_sms(dev) * per_sm;
    if (ctas > cap) ctas = cap;
    kernel<<<(unsigned)ctas, WARPS * 32, L::SMEM_BYTES, stream>>>(p);
    return (int)cudaGetLastError();
}

template <typename T>
int dispatch_rows(const StepParams& p, int dev, cudaStream_t stream) {
    switch (p.d) {
        case 2: return launch_rows_d<T, 2>(p, dev, stream);
        case 3: return launch_rows_d<T, 3>(p, dev, stream);
        case 4: return launch_rows_d<T, 4>(p, dev, stream);
        case 5: return launch_rows_d<T, 5>(p, dev, stream);
        default: return HK_ERR_UNSUPPORTED;
    }
}

}  // namespace
}  // namespace hk
