// hk_value.cuh — rollout value targets.
//
// calculate_value_using_reward_fn (hironaka/jax/util.py:261-284) as used by
// JAXTrainer.rollout_postprocess (hironaka/jax/jax_trainer.py:558-592): per game, the number of
// points at each of the T rollout steps (recovered from the observations: #(entries >= 0) / d -
// offset, :584) gives done flags, the single terminal reward, and
//   value[i] = sum_j clip(g^(j-i), -1, 1) * reward[j]  +  [game unfinished] * est * g^(T-1-i)
// with g = -discount for the unified (alternating) tree.  One warp per game; floating point,
// compared to the reference with a tolerance (the reference's own test uses isclose).
#pragma once
#include "hk_common.cuh"

namespace hk {

struct ValueParams {
    const float* obs;           // [B, T, W] (nullable when num_points is given)
    const int32_t* num_points;  // [B, T]   (nullable when obs is given)
    int32_t* num_points_out;    // [B, T]   (nullable)
    float* value;               // [B, T]
    long long B;
    int T, W, dimension, offset;
    float discount;
    int est_sign, reward_sign, unified;
};

constexpr int VALUE_MAX_T = 1024;

__global__ void __launch_bounds__(256) hk_value_targets_kernel(const ValueParams p) {
    extern __shared__ int32_t np_s[];  // [warps][T]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    int32_t* np = np_s + warp * p.T;
    const float g = p.unified ? -p.discount : p.discount;
    for (long long b = (long long)blockIdx.x * warps + warp; b < p.B; b += (long long)gridDim.x * warps) {
        // number of points per step
        if (p.num_points) {
            for (int t = lane; t < p.T; t += 32) np[t] = p.num_points[b * p.T + t];
        } else {
            for (int t = 0; t < p.T; ++t) {
                const float* row = p.obs + (b * p.T + t) * (long long)p.W;
                int c = 0;
                for (int w = lane; w < p.W; w += 32) c += (row[w] >= 0.0f) ? 1 : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                if (lane == 0) np[t] = c / p.dimension - p.offset;
            }
        }
        __syncwarp();
        if (p.num_points_out) {
            for (int t = lane; t < p.T; t += 32) p.num_points_out[b * p.T + t] = np[t];
        }
        const bool unfinished = np[p.T - 1] > 1;
        const int last = np[p.T - 1] < 1 ? 1 : np[p.T - 1];
        const float sgn = p.unified ? (((p.T + 1) & 1) ? -1.0f : 1.0f) : 1.0f;  // (-1)^(T+1)
        const float est = (1.0f / (float)last) * (float)p.est_sign * sgn;
        for (int i = lane; i < p.T; i += 32) {
            float v = 0.0f;
            for (int j = 0; j + 1 < p.T; ++j) {
                // reward_fn(next_done, done) = +-(done[j+1] & !done[j]); the last step has next_done = False
                const bool r = (np[j + 1] <= 1) && !(np[j] <= 1);
                if (r) {
                    float tab = powf(g, (float)(j - i));
                    tab = fminf(1.0f, fmaxf(-1.0f, tab));
                    v += tab * (float)p.reward_sign;
                }
            }
            if (unfinished) v += est * powf(g, (float)(p.T - 1 - i));
            p.value[b * p.T + i] = v;
        }
        __syncwarp();
    }
}

// Multi-binary coordinate vectors [B, d] (the host action form the reference passes around:
// decode_tensor output, hironaka/src/_fn.py:313-325) -> int32 bitmasks, bit k <=> coordinate k.
template <typename S>
__global__ void __launch_bounds__(256) hk_pack_coords_kernel(const S* coords, int32_t* mask, long long B, int d) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    uint32_t m = 0;
    for (int k = 0; k < d; ++k) m |= ((float)coords[b * d + k] > 0.5f) ? (1u << k) : 0u;
    mask[b] = (int32_t)m;
}

// Per-game overflow flags: overflow[b] = 1 iff some entry of game b reaches the threshold (>= ; > when strict) —
// TensorPoints.exceed_threshold per game (hironaka/core/tensor_points.py:57-63, whole-batch there) and the
// per-game rule of the gym environments (ListPoints.exceed_threshold, strict; hironaka_base.py:116-131).
// Dead rows hold a non-positive padding value, so looking at every entry is looking at the live ones.
// One warp per game, coalesced 4-byte reads, one vote.
template <typename T>
__global__ void __launch_bounds__(256) hk_overflow_kernel(const T* state, uint8_t* overflow, long long B, int W, float threshold,
                                                          int strict) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    for (long long b = (long long)blockIdx.x * warps + warp; b < B; b += (long long)gridDim.x * warps) {
        const T* x = state + b * W;
        bool over = false;
        for (int w = lane; w < W; w += 32) {
            const float v = (float)x[w];
            over = over || (strict ? (v > threshold) : (v >= threshold));
        }
        over = __any_sync(0xffffffffu, over);
        if (lane == 0) overflow[b] = over ? 1 : 0;
    }
}

// The action streams of the in-kernel random players, written out (hk_random_actions): the same draw as
// load_actions makes inside the step kernels.
__global__ void __launch_bounds__(256) hk_random_actions_kernel(int32_t* host_action_t, int32_t* axis_t, long long B, int d,
                                                                int T, unsigned long long seed, int step_offset) {
    const long long total = B * T;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long g = i % B;
        const int t = (int)(i / B);
        uint32_t r0, r1;
        philox4x32_10((uint32_t)g, (uint32_t)((unsigned long long)g >> 32), (uint32_t)(step_offset + t), 0u, (uint32_t)seed,
                      (uint32_t)(seed >> 32), r0, r1);
        if (host_action_t) host_action_t[i] = (int32_t)__umulhi(r0, (1u << d) - (uint32_t)d - 1u);
        if (axis_t) axis_t[i] = (int32_t)__umulhi(r1, (uint32_t)d);
    }
}

}  // namespace hk
