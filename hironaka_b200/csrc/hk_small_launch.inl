// hk_small_launch.inl — launch code of the thread-per-game kernel, included by hk_small_*.cu
#include "hk_launch.cuh"
#include "hk_small.cuh"

namespace hk {
namespace {

// ring geometry of the thread-per-game kernel per shape: (warps per CTA, stages per warp).
// Since unchanged games are no longer written back, the steady state of a rollout is bound by the
// instruction latency of each warp's tile loop rather than by bytes in flight, and warps per SM
// count for more than stages per warp (tools/tune_small.cu on B200, C2 workload, us/step:
// 8x1 = 72.9, 4x1 = 73.8, 4x2 = 73.8, 5x2 = 82.9, 4x3 = 84.5; with every game written, 4x3 = 94.8).
template <int N, int D, bool OBS>
struct SmallTune {
    static constexpr int WARPS = OBS ? 4 : 8;
    // step + features is issue- and latency-bound in the first steps of a rollout (many live rows):
    // one stage per warp and 12 warps per SM beat two stages and 8 warps (143.6 vs 152.7 us/step at C2)
    static constexpr int STAGES = 1;
};

template <typename T, int N, int D, bool OBS, bool POLICY, int WARPS, int STAGES, bool PACKED = true>
int launch_small_geom(const StepParams& p, int dev, cudaStream_t stream) {
    using L = SmallLayout<N, D, OBS, WARPS, STAGES>;
    static KernelFacts facts;
    auto kernel = hk_small_kernel<T, N, D, OBS, POLICY, WARPS, STAGES, PACKED>;
    cudaError_t err = cudaSuccess;
    const int threads = WARPS * 32;
    const int per_sm = kernel_ctas_per_sm(kernel, facts, dev, threads, L::SMEM_BYTES, &err);
    if (err != cudaSuccess) return (int)err;
    const long long ntiles = (p.B + 31) / 32;
    long long ctas = (ntiles + WARPS - 1) / WARPS;
    const long long cap = (long long)device_sms(dev) * per_sm;  // persistent: one wave
    if (ctas > cap) ctas = cap;
    if (use_pdl()) {
        // programmatic stream serialisation: this grid's CTAs may be scheduled while the previous
        // launch of the stream drains; the kernel waits (griddepcontrol.wait) before any global access
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)ctas);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = L::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return (int)cudaLaunchKernelEx(&cfg, kernel, p);
    }
    kernel<<<(unsigned)ctas, threads, L::SMEM_BYTES, stream>>>(p);
    return (int)cudaGetLastError();
}

template <typename T, int N, int D, bool OBS, bool POLICY>
int launch_small_shape(const StepParams& p, int dev, cudaStream_t stream) {
    // A one-launch rollout (T > 1) reads and writes the state once per T steps, so it is bound by
    // issue rate, not by bytes in flight: one stage per warp and twice the warps
    // (tools/tune_small.cu: 8x1 = 0.648 ms, 4x1 = 0.661, 4x2 = 0.733, 4x3 = 0.808 per 20-step rollout).
    if constexpr (!OBS) {
        // (its own instantiation without the packed tiers where those exist: int32 state, no fixed players)
        if (p.T > 1) return launch_small_geom<T, N, D, false, POLICY, 8, 1, Elem<T>::is_float || POLICY>(p, dev, stream);
    }
    return launch_small_geom<T, N, D, OBS, POLICY, SmallTune<N, D, OBS>::WARPS, SmallTune<N, D, OBS>::STAGES>(p, dev, stream);
}

// The fixed players (Zeillinger host etc.) live in their own instantiations so that the kernels
// of the ordinary step do not carry that code in their hot loops.
template <typename T, bool OBS>
int dispatch_small(const StepParams& p, int dev, cudaStream_t stream) {
    const bool policy = p.flags & (HK_F_HOST_ALL_COORD | HK_F_HOST_ZEILLINGER | HK_F_AGENT_FIRST | HK_F_AGENT_LAST);
    if (p.d == 3) {
        if (p.N == 20) return policy ? launch_small_shape<T, 20, 3, OBS, true>(p, dev, stream) : launch_small_shape<T, 20, 3, OBS, false>(p, dev, stream);
        if (p.N == 10) return policy ? launch_small_shape<T, 10, 3, OBS, true>(p, dev, stream) : launch_small_shape<T, 10, 3, OBS, false>(p, dev, stream);
        if (p.N == 5) return policy ? launch_small_shape<T, 5, 3, OBS, true>(p, dev, stream) : launch_small_shape<T, 5, 3, OBS, false>(p, dev, stream);
    }
    return HK_ERR_UNSUPPORTED;
}

}  // namespace
}  // namespace hk
