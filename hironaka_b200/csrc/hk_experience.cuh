// hk_experience.cuh — order-preserving experience writer into circular replay buffers.
//
// Replaces the boolean-mask filtering of FusedGame.step (hironaka/trainer/fused_game.py:82-99:
// obs[~done], actions[~done], ... — a device->host sync per tensor in the reference) followed by
// ReplayBuffer.add (hironaka/trainer/replay_buffer.py:63-127, wrap-around at :116-124).  Rows whose
// `skip` byte is 0 are appended in batch order at (pos + rank) mod capacity; `pos` and `full` live
// on the device, so a step can be appended without the host ever learning how many rows it kept.
// Three stream-ordered launches: per-block keep counts, a one-block exclusive scan (which also
// advances pos/full), and the scatter (one warp per row, coalesced row copies).
#pragma once
#include "hk_common.cuh"

namespace hk {

constexpr int EXP_ROWS_PER_BLOCK = 2048;
constexpr int EXP_THREADS = 256;

struct ExpParams {
    const uint8_t* skip;   // [B] 1 = drop
    const float* obs;      // [B, ow] (nullable)
    const float* next_obs; // [B, ow]
    const float* coords;   // [B, cw] (nullable)
    const float* next_coords;
    const int32_t* action; // [B]
    const float* reward;   // [B]
    const uint8_t* done;   // [B]
    float* buf_obs;
    float* buf_next_obs;
    float* buf_coords;
    float* buf_next_coords;
    int32_t* buf_action;
    float* buf_reward;
    uint8_t* buf_done;
    long long capacity;
    long long* pos;    // device, in/out
    int32_t* full;     // device, in/out
    int32_t* appended; // device out (nullable)
    int32_t* scratch;  // [nblocks + 4]: block offsets, then base pos (2 words) and total
    long long B;
    int ow, cw, nblocks;
};

__global__ void __launch_bounds__(EXP_THREADS) hk_exp_count_kernel(const ExpParams p) {
    const long long base = (long long)blockIdx.x * EXP_ROWS_PER_BLOCK;
    int c = 0;
    for (int i = threadIdx.x; i < EXP_ROWS_PER_BLOCK; i += EXP_THREADS) {
        const long long r = base + i;
        c += (r < p.B && p.skip[r] == 0) ? 1 : 0;
    }
    __shared__ int red[EXP_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < EXP_THREADS / 32; ++w) t += red[w];
        p.scratch[blockIdx.x] = t;
    }
}

// one block: exclusive scan of the block counts in place; publishes the base position and the
// total, then advances the device-resident pos / full (replay_buffer.py:126-127)
__global__ void __launch_bounds__(1024) hk_exp_scan_kernel(const ExpParams p) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int start = 0; start < p.nblocks; start += 1024) {
        const int i = start + threadIdx.x;
        const int v = (i < p.nblocks) ? p.scratch[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += u;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_tot[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += u;
            }
            warp_tot[threadIdx.x] = w;  // inclusive over warps
        }
        __syncthreads();
        const int warp_excl = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0;
        const int carry = carry_s;
        if (i < p.nblocks) p.scratch[i] = carry + warp_excl + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_excl + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int total = carry_s;
        const long long pos0 = *p.pos;
        p.scratch[p.nblocks] = (int32_t)(pos0 & 0xffffffffll);
        p.scratch[p.nblocks + 1] = (int32_t)(pos0 >> 32);
        p.scratch[p.nblocks + 2] = total;
        if (p.appended) *p.appended = total;
        if (pos0 + total >= p.capacity) *p.full = 1;
        *p.pos = (pos0 + total) % p.capacity;
    }
}

__device__ __forceinline__ void copy_row(float* dst, const float* src, int w, int lane) {
    for (int c = lane; c < w; c += 32) dst[c] = src[c];
}

__global__ void __launch_bounds__(EXP_THREADS) hk_exp_scatter_kernel(const ExpParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WARPS = EXP_THREADS / 32;
    constexpr int ROWS_PER_WARP = EXP_ROWS_PER_BLOCK / WARPS;  // contiguous rows per warp keep the order
    const long long base = (long long)blockIdx.x * EXP_ROWS_PER_BLOCK;
    const long long pos0 = ((long long)(uint32_t)p.scratch[p.nblocks]) | ((long long)p.scratch[p.nblocks + 1] << 32);
    // ranks inside the block: count the kept rows of the preceding warps' row ranges
    __shared__ int warp_cnt[WARPS];
    const long long wbase = base + (long long)warp * ROWS_PER_WARP;
    int c = 0;
    for (int i = lane; i < ROWS_PER_WARP; i += 32) {
        const long long r = wbase + i;
        c += (r < p.B && p.skip[r] == 0) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) warp_cnt[warp] = c;
    __syncthreads();
    long long rank = p.scratch[blockIdx.x];
    for (int w = 0; w < warp; ++w) rank += warp_cnt[w];
    for (int i0 = 0; i0 < ROWS_PER_WARP; i0 += 32) {
        const long long r = wbase + i0 + lane;
        const bool keep = (r < p.B) && (p.skip[r] == 0);
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const long long dst = (pos0 + rank + __popc(bal & ((1u << lane) - 1u))) % p.capacity;
            if (p.buf_action) p.buf_action[dst] = p.action[r];
            if (p.buf_reward) p.buf_reward[dst] = p.reward[r];
            if (p.buf_done) p.buf_done[dst] = p.done[r];
        }
        // the wide rows: the whole warp copies one kept row at a time (coalesced)
        uint32_t m = bal;
        while (m) {
            const int l = __ffs((int)m) - 1;
            m &= m - 1;
            const long long rr = wbase + i0 + l;
            const long long dst = (pos0 + rank + __popc(bal & ((1u << l) - 1u))) % p.capacity;
            if (p.buf_obs) copy_row(p.buf_obs + dst * p.ow, p.obs + rr * p.ow, p.ow, lane);
            if (p.buf_next_obs) copy_row(p.buf_next_obs + dst * p.ow, p.next_obs + rr * p.ow, p.ow, lane);
            if (p.buf_coords) copy_row(p.buf_coords + dst * p.cw, p.coords + rr * p.cw, p.cw, lane);
            if (p.buf_next_coords) copy_row(p.buf_next_coords + dst * p.cw, p.next_coords + rr * p.cw, p.cw, lane);
        }
        rank += __popc(bal);
    }
}

}  // namespace hk
