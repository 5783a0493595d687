// hk_sched_launch.inl — launch code of the census-scheduled kernel, included by hk_sched_*.cu
#include "hk_launch.cuh"
#include "hk_sched.cuh"

namespace hk {
namespace {

template <typename T, int N, int D, int WARPS, int STAGES, bool OBS = false, int MINB = 1>
int launch_sched_geom(const StepParams& p, int dev, cudaStream_t stream) {
    using L = SchedLayout<N, D, WARPS, STAGES, OBS>;
    static KernelFacts facts;
    auto kernel = hk_sched_kernel<T, N, D, WARPS, STAGES, OBS, MINB>;
    cudaError_t err = cudaSuccess;
    const int threads = WARPS * 32;
    const int per_sm = kernel_ctas_per_sm(kernel, facts, dev, threads, L::SMEM_BYTES, &err);
    if (err != cudaSuccess) return (int)err;
    const long long ntiles = (p.B + 31) / 32;
    long long ctas = (ntiles + WARPS - 1) / WARPS;
    const long long cap = (long long)device_sms(dev) * per_sm;  // persistent: one wave
    if (ctas > cap) ctas = cap;
    kernel<<<(unsigned)ctas, threads, L::SMEM_BYTES, stream>>>(p);
    return (int)cudaGetLastError();
}

template <typename T, int N, int D, bool OBS>
int launch_sched_shape(const StepParams& p, int dev, cudaStream_t stream) {
    // with the observation tile: 4 warps x 1 stage (12 warps per SM), as the step + features kernel of K-small
    // (held to 3 CTAs per SM: with the packed tiers ptxas would take 242 registers and leave room for two)
    if constexpr (OBS) return launch_sched_geom<T, N, D, 4, 1, true, Elem<T>::is_float ? 1 : 3>(p, dev, stream);
    // geometry (warps per CTA x stages per warp): see DESIGN.md "K-sched"; hk_debug_set_sched_geometry switches it
    else if constexpr (!Elem<T>::is_float) {
        // int32 state steps on packed rows (hk_small.cuh, tier_packed): few registers in the hot path, so the kernel is
        // held to 128 registers (4 CTAs of 4 warps per SM, one stage each: 16 warps per SM against 12 with two stages)
        switch (sched_geometry()) {
            case 1: return launch_sched_geom<T, N, D, 8, 1>(p, dev, stream);
            // (measured: 3 warps x 6 CTAs and 6 warps x 3 CTAs, 18 warps per SM: ptxas settles on 96 registers with 136 bytes
            // of stack, 0.0498 - 0.0511 ms per step against 0.0448)
            // (timing experiment with the exact 8/12/16-row tiers compiled out — 11.2k instructions instead of 15.2k, 96
            // registers with 52 bytes of spills at 20 warps per SM: first steps 135 107 109 107 us, no better than 16 warps
            // with the full kernel; neither code size nor occupancy is what holds the first steps)
            default: return launch_sched_geom<T, N, D, 4, 1, false, 4>(p, dev, stream);
        }
    }
    else {
        switch (sched_geometry()) {
            case 1: return launch_sched_geom<T, N, D, 8, 1>(p, dev, stream);
            default: return launch_sched_geom<T, N, D, 4, 2>(p, dev, stream);
        }
    }
}

template <typename T, bool OBS>
int dispatch_sched(const StepParams& p, int dev, cudaStream_t stream) {
    if (p.d == 3) {
        if (p.N == 20) return launch_sched_shape<T, 20, 3, OBS>(p, dev, stream);
        if (p.N == 10) return launch_sched_shape<T, 10, 3, OBS>(p, dev, stream);
        if (p.N == 5) return launch_sched_shape<T, 5, 3, OBS>(p, dev, stream);
    }
    return HK_ERR_UNSUPPORTED;
}

}  // namespace
}  // namespace hk
