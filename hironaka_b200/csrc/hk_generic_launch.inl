// hk_generic_launch.inl — launch code of the warp-per-game kernel, included by hk_generic_*.cu
#include <type_traits>

#include "hk_generic.cuh"
#include "hk_launch.cuh"

namespace hk {
namespace {

template <typename T, int D, bool OBS, int RT, int DEPTH, bool HOT, bool CENSUS = false>
int launch_generic_hot(const StepParams& p, int dev, cudaStream_t stream) {
    auto kernel = hk_generic_kernel<T, D, OBS, RT, DEPTH, HOT, CENSUS>;
    const int W = p.N * D;
    const int Wpad = (W + 3) & ~3;
    const int R = (p.N + 31) / 32;
    // DEPTH + 1 state buffers (+ features) + live-mask words + the compact list of live rows (N + 1 rows)
    const int slot_words = Wpad * (DEPTH + 1 + (OBS ? 1 : 0)) + ((R + 3) & ~3) +
                           (((p.N + 1) * generic_compact_stride(D) + 3) & ~3);
    int warps = 8;
    while (warps > 1 && (size_t)warps * slot_words * 4 > 160 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * slot_words * 4;
    // launch facts depend on N through the shared-memory size: cache the last one per device
    struct Facts {
        std::atomic<size_t> smem_set{0};
        std::atomic<long long> key{-1};
        std::atomic<int> per_sm{0};
    };
    static Facts facts[kMaxDevices];
    Facts& fc = facts[dev];
    cudaError_t err = cudaSuccess;
    if (smem > fc.smem_set.load(std::memory_order_acquire)) {
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        fc.smem_set.store(smem, std::memory_order_release);
    }
    const long long key = ((long long)smem << 8) | warps;
    int per_sm = fc.per_sm.load(std::memory_order_acquire);
    if (fc.key.load(std::memory_order_acquire) != key || per_sm <= 0) {
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem);
        if (err != cudaSuccess) return (int)err;
        if (per_sm < 1) per_sm = 1;
        fc.per_sm.store(per_sm, std::memory_order_release);
        fc.key.store(key, std::memory_order_release);
    }
    long long ctas = (p.B + warps - 1) / warps;
    const long long cap = (long long)device_sms(dev) * per_sm;
    if (ctas > cap) ctas = cap;
    kernel<<<(unsigned)ctas, warps * 32, smem, stream>>>(p, warps, slot_words);
    return (int)cudaGetLastError();
}

// One game ahead in flight per warp.  Three ahead (DEPTH = 3, four buffers) measured slower at C5
// (0.232 vs 0.224 ms per step, tail 0.148 vs 0.138): the kernel is bound by instruction issue, not by
// the latency of its loads.
template <typename T, int D, bool OBS, int RT>
int launch_generic_rt(const StepParams& p, int dev, cudaStream_t stream) {
    // the plain random-play step of an int32 state has an instantiation with its op / flag words folded in
    if constexpr (!OBS && RT > 0 && std::is_same<T, int32_t>::value) {
        const bool hot = p.T == 1 && p.ops == GENERIC_HOT_OPS && p.flags == GENERIC_HOT_FLAGS && !p.host_out &&
                         p.host_action && p.axis;
        if (hot) return p.census ? launch_generic_hot<T, D, OBS, RT, 1, true, true>(p, dev, stream)
                                 : launch_generic_hot<T, D, OBS, RT, 1, true>(p, dev, stream);
    }
    if constexpr (!OBS) {
        if (p.census) return launch_generic_hot<T, D, OBS, RT, 1, false, true>(p, dev, stream);
    }
    return launch_generic_hot<T, D, OBS, RT, 1, false>(p, dev, stream);
}

// rows-per-lane specialisations exist for the common dimensions; everything else takes run-time loops
template <typename T, int D, bool OBS>
int launch_generic(const StepParams& p, int dev, cudaStream_t stream) {
    if constexpr (D >= 2 && D <= 5) {
        if (p.N <= 32) return launch_generic_rt<T, D, OBS, 1>(p, dev, stream);
        if (p.N <= 64) return launch_generic_rt<T, D, OBS, 2>(p, dev, stream);
    }
    return launch_generic_rt<T, D, OBS, 0>(p, dev, stream);
}

template <typename T, bool OBS>
int dispatch_generic(const StepParams& p, int dev, cudaStream_t stream) {
    switch (p.d) {
        case 1: return launch_generic<T, 1, OBS>(p, dev, stream);
        case 2: return launch_generic<T, 2, OBS>(p, dev, stream);
        case 3: return launch_generic<T, 3, OBS>(p, dev, stream);
        case 4: return launch_generic<T, 4, OBS>(p, dev, stream);
        case 5: return launch_generic<T, 5, OBS>(p, dev, stream);
        case 6: return launch_generic<T, 6, OBS>(p, dev, stream);
        case 7: return launch_generic<T, 7, OBS>(p, dev, stream);
        case 8: return launch_generic<T, 8, OBS>(p, dev, stream);
        case 9: return launch_generic<T, 9, OBS>(p, dev, stream);
        case 10: return launch_generic<T, 10, OBS>(p, dev, stream);
        default: return HK_ERR_UNSUPPORTED;
    }
}

}  // namespace
}  // namespace hk
