// hk_sched.cuh — census-scheduled thread-per-game step kernel (in-place single steps of small games).
//
// The tile-ring kernel (hk_small.cuh) reads every game every step and lets a tile's BUSIEST game pick the
// tier for all 32.  In a rollout that wastes most of the work: a game that has ended sits at a fixed point
// (a lone point at the origin: 52 % of the games after 5 random-play steps at (20,3), 95 % after 10), and the
// live counts inside a tile are heavy-tailed (mean 6 rows, tile maximum 11 at the root).  This kernel keeps
// a CENSUS — one byte per game, written by the previous step — and schedules by it:
//   * games at rest are not loaded at all: their done = 1 / reward = 0 come from the census byte;
//   * the other games of a warp's tiles are ordered by live count (counting sort over six tier classes)
//     and cut into chunks of 32, so that a chunk's tier fits all of its games;
//   * each chunk is gathered game by game: where a game is a whole number of 16-byte words ((20,3): 240 B)
//     every lane issues ONE TMA bulk copy (cp.async.bulk, SASS UBLKCP) for its own game onto the stage's
//     mbarrier, and a changed game goes back with one bulk store per lane; the other shapes ((10,3), (5,3),
//     unaligned pointers) are copied with cp.async in 8/4-byte pieces, 15 lanes per game.  The chunk is
//     processed by the same per-game code as the tile-ring kernel (small_process_tile).
// Work and DRAM traffic follow the number of games still in play instead of the batch size.
//
// Census byte (HK census, include/hironaka_b200.h): 0 = unknown (the game is read and counted);
// 1..127 = live rows of a game in play; 0x80 | c | (o ? 2 : 0) = ended game with c <= 1 live rows, dead rows
// normalised, o = at rest for every op (no row, or the lone row at the origin).
#pragma once
#include "hk_small.cuh"

namespace hk {

constexpr int SCHED_ROUND_TILES = 32;  // tiles (of 32 games) a warp schedules together
#ifndef SCHED_PASS_UNROLL
#define SCHED_PASS_UNROLL 2
#endif
constexpr int kSchedPassUnroll = SCHED_PASS_UNROLL;
#ifndef SCHED_TILE_GROUP
#define SCHED_TILE_GROUP 0  // 0: a warp owns one contiguous run of tiles; G > 0: runs of G tiles, dealt round-robin to the warps
#endif
#ifndef SCHED_DIRECT
#define SCHED_DIRECT 1  // packed tiers release the stage after the gather and store changed rows straight to global memory
#endif
#ifndef SCHED_SPECULATE
#define SCHED_SPECULATE 0  // (measured, not kept: no gain on the first steps, 3-5 % slower in the middle of a rollout) the first tile of a round is requested before the census pass has finished (see pass 1)
#endif
#ifndef SCHED_NATURAL_NUM
#define SCHED_NATURAL_NUM 1  // a round is stepped tile by tile in natural order while at least NUM / DEN of
#define SCHED_NATURAL_DEN 2  // its games are in play (tools/time_census.py)
#endif

template <int N, int D, int WARPS, int STAGES, bool OBS = false>
struct SchedLayout {
    static constexpr int W = N * D;
    static constexpr int STAGE_WORDS = 32 * W;
    static constexpr int OBS_WORDS = OBS ? 32 * (W + D) : 0;  // the observation rows of one chunk
    // per warp: STAGES game stages, the class bytes of a round (one per game), the sorted order (u16 per game)
    static constexpr int CLS_WORDS = SCHED_ROUND_TILES * 32 / 4;
    static constexpr int ORDER_WORDS = SCHED_ROUND_TILES * 32 / 2;
    static constexpr int WARP_WORDS = STAGES * STAGE_WORDS + CLS_WORDS + ORDER_WORDS + OBS_WORDS;
    static constexpr int BAR_BYTES = 256;  // mbarriers first (WARPS * STAGES of them), then the warps' areas
    static_assert(WARPS * STAGES * 8 <= BAR_BYTES, "mbarrier area too small");
    static constexpr size_t SMEM_BYTES = BAR_BYTES + (size_t)WARPS * WARP_WORDS * 4;
};

__device__ __forceinline__ int census_class(uint32_t v) {
    if (v == 0) return 6;          // unknown: counted after the load
    if (v & 0x80u) return 1;       // ended, but not at rest under this launch's ops
    return v <= 2 ? 1 : (v <= 4 ? 2 : (v <= 8 ? 3 : (v <= 12 ? 4 : (v <= 16 ? 5 : 6))));
}

// OBS: the instantiation that also builds the observation features of the new state (p.obs, sorted modes only):
// a game at rest has a CONSTANT observation — its lone point, at the origin, sorts first, the rest is padding — so
// it is written from the census byte like the other outputs, and only the games in play run the feature code.
template <typename T, int N, int D, int WARPS, int STAGES, bool OBS = false, int MINB = 1>
__global__ void __launch_bounds__(WARPS * 32, MINB) hk_sched_kernel(const StepParams p) {
    using L = SchedLayout<N, D, WARPS, STAGES, OBS>;
    constexpr int W = L::W;
    constexpr int CHW = (W % 4 == 0) ? 4 : ((W % 2 == 0) ? 2 : 1);  // words per copy piece
    constexpr int LPG = W / CHW;                                     // lanes per game
    static_assert(LPG <= 32, "a game is copied by at most one warp instruction");
    constexpr int GPI = 32 / LPG;                                    // games per copy instruction
    constexpr int COPY_ITERS = (32 + GPI - 1) / GPI;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw) + warp * STAGES;
    uint32_t* wbase = reinterpret_cast<uint32_t*>(smem_raw + L::BAR_BYTES) + (size_t)warp * L::WARP_WORDS;
    uint32_t* stages = wbase;
    uint8_t* clsb = reinterpret_cast<uint8_t*>(wbase + STAGES * L::STAGE_WORDS);
    uint16_t* order = reinterpret_cast<uint16_t*>(wbase + STAGES * L::STAGE_WORDS + L::CLS_WORDS);
    float* obs_tile = reinterpret_cast<float*>(wbase + STAGES * L::STAGE_WORDS + L::CLS_WORDS + L::ORDER_WORDS);
    const int OW = W + (p.obs_coord ? D : 0);
    const bool obs_bulk = OBS && aligned16(p.obs) && ((OW & 3) == 0);

    const long long B = p.B;
    const long long ntiles = (B + 31) >> 5;
    const long long gw = (long long)blockIdx.x * WARPS + warp;
    const long long nw = (long long)gridDim.x * WARPS;
    uint32_t* gst = reinterpret_cast<uint32_t*>(p.out);  // in place: p.out == p.in
    const bool piece_ok = ((reinterpret_cast<uintptr_t>(p.out)) & (CHW * 4 - 1)) == 0;
    // one TMA bulk copy per game when a game is a whole number of 16-byte words and the state is 16-byte aligned
    const bool bulk = ((W * 4) % 16 == 0) && aligned16(p.out);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t phase_bits = 0;  // bit s = parity to wait for on stage s
    const uint32_t lt = (1u << lane) - 1u;
    // at rest under THIS launch: nothing to do for an origin / empty game; a frozen ended game also rests when
    // neither reposition nor rescale would move its lone point
    const bool frozen_rest = (p.flags & HK_F_FREEZE_ENDED) && !(p.ops & (HK_OP_REPOSITION | HK_OP_RESCALE));
    const float rest_reward = (p.flags & HK_F_ROLE_AGENT) ? -0.0f : 0.0f;
    // piece lane -> (game slot within the instruction, piece within the game)
    const int sub = lane / LPG, piece = lane - sub * LPG;
    int settled_total = 0;

    // A warp owns a CONTIGUOUS run of tiles (its games share DRAM pages when they are gathered one by one),
    // scheduled in rounds of SCHED_ROUND_TILES.
#if SCHED_TILE_GROUP == 0
    const long long tpw = (ntiles + nw - 1) / nw;
    const long long t_begin = gw * tpw;
    const long long t_end = (t_begin + tpw < ntiles) ? (t_begin + tpw) : ntiles;
    auto tile_of = [&](long long j) -> long long { return t_begin + j; };  // the warp's j-th tile
#else
    constexpr int TG = SCHED_TILE_GROUP;
    const long long t_end = ntiles;
    auto tile_of = [&](long long j) -> long long { return ((j / TG) * nw + gw) * TG + (j % TG); };
#endif
    for (long long j0 = 0; tile_of(j0) < t_end; j0 += SCHED_ROUND_TILES) {
        // ---- census of this round's tiles: outputs of the games at rest, class bytes of the others ----
#if SCHED_TILE_GROUP == 0
        const int nk = (int)((t_end - t_begin - j0 < SCHED_ROUND_TILES) ? (t_end - t_begin - j0) : SCHED_ROUND_TILES);
#else
        int nk = 0;  // tiles of this round that exist: whole runs, then the part of the run that crosses the end
#pragma unroll
        for (int q = 0; q < SCHED_ROUND_TILES / TG; ++q) {
            const long long left = t_end - tile_of(j0 + q * TG);
            nk += (left >= TG) ? TG : (left > 0 ? (int)left : 0);
        }
#endif
        uint32_t tilemask = 0, classes = 0;
        int inplay = 0, ngames = 0;
        bool spec = false;  // the round's first tile has been requested ahead of the schedule
        // The round's census bytes are staged in shared memory (the class array doubles as the landing zone):
        // cp.async in 4-byte pieces, 8 lanes per tile, all requests in flight at once, then ONE rolled loop
        // over the tiles (an unrolled register version was 16 KB of straight-line code in a kernel whose hot
        // loops compete for the instruction cache).
        {
            const uint8_t* cbase = p.census + (tile_of(j0) << 5);
            const long long cbytes = ((B - (tile_of(j0) << 5)) < (long long)nk * 32) ? (B - (tile_of(j0) << 5)) : (long long)nk * 32;
            if (SCHED_TILE_GROUP == 0 && ((reinterpret_cast<uintptr_t>(cbase) & 3u) == 0)) {
                for (int o = lane * 4; o < nk * 32; o += 128) {
                    if (o + 4 <= cbytes) cp_async_4(clsb + o, cbase + o);
                    else {
                        for (int q = 0; q < 4; ++q) clsb[o + q] = (o + q < cbytes) ? __ldg(cbase + o + q) : (uint8_t)0;
                    }
                }
                cp_async_commit();
                cp_async_wait<0>();
            } else {
                for (int k = 0; k < nk; ++k) {
                    const long long g = (tile_of(j0 + k) << 5) + lane;
                    clsb[k * 32 + lane] = (g < B) ? __ldg(p.census + g) : (uint8_t)0;
                }
            }
            __syncwarp();
        }
        // pass 1: class of every game (0 = at rest or no such game), what is in play; without the fused observation
        // also the outputs of the games at rest, from their census bytes (with it they are a second pass, below, that
        // knows which tiles are about to be stepped whole).  Late in a rollout this loop IS the step (a warp walks its
        // tiles, finds a handful of games in play): unrolled by two so that the tiles' dependency chains interleave.
#pragma unroll kSchedPassUnroll
        for (int k = 0; k < nk; ++k) {
            const long long g = (tile_of(j0 + k) << 5) + lane;
            const bool valid = g < B;
            const uint32_t v = clsb[k * 32 + lane];
            const bool rest = valid && (v & 0x80u) && ((v & 2u) || frozen_rest);
            const int c = (!valid || rest) ? 0 : census_class(v);
            clsb[k * 32 + lane] = (uint8_t)(c | (rest ? 0x80 : 0) | ((v & 1u) << 6));  // class | at rest | its live count
            const uint32_t play = __ballot_sync(0xffffffffu, c != 0);
            if (play) tilemask |= 1u << k;
            inplay += __popc(play);
            const uint32_t vbal = __ballot_sync(0xffffffffu, valid);
            ngames += __popc(vbal);
            classes |= 1u << c;
            if constexpr (SCHED_SPECULATE) {
                // The round's first tile predicts the round: when at least half of ITS games are in play the tile is
                // requested now, as the natural order would, and the rest of this pass runs in the shadow of that
                // load (early in a rollout the pass was a bubble of ~3 us in front of every warp's first tile).  A
                // wrong guess (the round turns out to be sorted) costs one tile of traffic; its arrival is awaited
                // below before the stage is used again.
                if (k == 0 && bulk && play && __popc(play) * 2 >= __popc(vbal)) {
                    spec = true;
                    const long long first = g - lane;
                    bulk_wait_read<0>();
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_expect_tx(&bar[0], (uint32_t)__popc(vbal) * (uint32_t)(W * 4));
                        bulk_load(stages, gst + first * W, (uint32_t)__popc(vbal) * (uint32_t)(W * 4), &bar[0]);
                    }
                }
            }
            if constexpr (!OBS) {
                if (rest) {
                    if (p.done) p.done[g] = 1;
                    if (p.reward) p.reward[g] = rest_reward;
                    if (p.num_points) p.num_points[g] = (int32_t)(v & 1u);
                }
                const uint32_t restmask = __ballot_sync(0xffffffffu, rest);
                settled_total += __popc(restmask);
                // the tile's word of the done mask: the games at rest now, the games that finish in this step later
                if (p.done_bits && lane == 0) p.done_bits[tile_of(j0 + k)] = restmask;
            }
        }
        classes = __reduce_or_sync(0xffffffffu, classes) & ~1u;
        // While most games are still in play the tiles are stepped whole, in natural order: one bulk copy per
        // tile and coalesced actions / outputs beat the per-game gather, and there is little to skip.  Later the
        // games in play are sorted by class and gathered one by one.
        // (with the fused observation the per-game stores of the sorted order cost more: natural order down to a quarter)
        const bool natural = OBS ? (inplay * 4 >= ngames) : (inplay * SCHED_NATURAL_DEN >= ngames * SCHED_NATURAL_NUM);
        __syncwarp();
        // ---- counting sort of the games in play by class: order[] lists them (tile slot * 32 + lane) ----
        int total = 0;
        if (!natural && inplay <= 32) {  // one chunk whatever the order: a single compaction pass
#pragma unroll kSchedPassUnroll
            for (int k = 0; k < nk; ++k) {
                if (!((tilemask >> k) & 1u)) continue;
                const bool m = (clsb[k * 32 + lane] & 7u) != 0u;
                const uint32_t bal = __ballot_sync(0xffffffffu, m);
                if (m) order[total + __popc(bal & lt)] = (uint16_t)(k * 32 + lane);
                total += __popc(bal);
            }
        }
        for (int q = 1; q <= 6 && !natural && inplay > 32; ++q) {
            if (!((classes >> q) & 1u)) continue;
            for (int k = 0; k < nk; ++k) {
                if (!((tilemask >> k) & 1u)) continue;
                const bool m = (clsb[k * 32 + lane] & 7u) == (uint32_t)q;
                const uint32_t bal = __ballot_sync(0xffffffffu, m);
                if (m) order[total + __popc(bal & lt)] = (uint16_t)(k * 32 + lane);
                total += __popc(bal);
            }
        }
        __syncwarp();
        const int nchunks = natural ? __popc(tilemask) : ((total + 31) >> 5);
        uint32_t tiles_left = tilemask;  // natural order: the tiles still to be requested

        // lane's game of chunk v (natural order: the next tile with a game in play; its games at rest ride along
        // in the tile copy but are not stepped: their outputs were written above)
        auto chunk_game = [&](int v, long long& g, bool& valid) {
            if (natural) {
                const int k = __ffs((int)tiles_left) - 1;
                tiles_left &= tiles_left - 1;
                g = (tile_of(j0 + k) << 5) + lane;
                valid = (g < B) && ((clsb[k * 32 + lane] & 7u) != 0);
                return;
            }
            const int idx = v * 32 + lane;
            valid = idx < total;
            const int lid = valid ? (int)order[idx] : 0;
            g = (tile_of(j0 + (lid >> 5)) << 5) + (lid & 31);
        };
        // gather the chunk's games into a stage: game slot s at s * W words
        auto gather = [&](long long g, bool valid, uint32_t* stage, int sidx) {
            if (natural) {  // the whole tile: games g - lane .. of which (B - first) may be fewer than 32
                const long long first = g - lane;
                const int cnt = (int)((B - first < 32) ? (B - first) : 32);
                if (bulk) {
                    bulk_wait_read<0>();
                    // the lanes wrote this stage with st.shared (results, scratch) and it may not have been stored: order
                    // those generic-proxy writes before the async-proxy write of the refill
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_expect_tx(&bar[sidx], (uint32_t)cnt * (uint32_t)(W * 4));
                        bulk_load(stage, gst + first * W, (uint32_t)cnt * (uint32_t)(W * 4), &bar[sidx]);
                    }
                    return;
                }
                if (piece_ok) {
                    for (int w = lane * CHW; w < cnt * W; w += 32 * CHW) {
                        if constexpr (CHW == 4) cp_async_16(stage + w, gst + first * W + w);
                        else if constexpr (CHW == 2) cp_async_8(stage + w, gst + first * W + w);
                        else cp_async_4(stage + w, gst + first * W + w);
                    }
                } else {
                    for (int w = lane; w < cnt * W; w += 32) cp_async_4(stage + w, gst + first * W + w);
                }
                cp_async_commit();
                return;
            }
            if (bulk) {
                // the lane's own earlier bulk store from this stage must have finished READING it
                bulk_wait_read<0>();
                fence_async_smem();  // (as above: earlier st.shared writes of this stage before the async-proxy refill)
                const int nvalid = __popc(__ballot_sync(0xffffffffu, valid));
                if (lane == 0) mbar_expect_tx(&bar[sidx], (uint32_t)nvalid * (uint32_t)(W * 4));
                __syncwarp();
                if (valid) bulk_load(stage + lane * W, gst + g * W, (uint32_t)(W * 4), &bar[sidx]);
                return;
            }
            if (piece_ok) {
#pragma unroll
                for (int it = 0; it < COPY_ITERS; ++it) {
                    const int s = it * GPI + sub;
                    const long long gs = __shfl_sync(0xffffffffu, g, s & 31);
                    const bool vs = __shfl_sync(0xffffffffu, valid ? 1 : 0, s & 31) != 0;
                    if (sub < GPI && s < 32 && vs) {
                        uint32_t* dst = stage + s * W + piece * CHW;
                        const uint32_t* src = gst + gs * W + piece * CHW;
                        if constexpr (CHW == 4) cp_async_16(dst, src);
                        else if constexpr (CHW == 2) cp_async_8(dst, src);
                        else cp_async_4(dst, src);
                    }
                }
            } else {  // state pointer aligned to 4 bytes only: word copies
                for (int s = 0; s < 32; ++s) {
                    const long long gs = __shfl_sync(0xffffffffu, g, s);
                    const bool vs = __shfl_sync(0xffffffffu, valid ? 1 : 0, s) != 0;
                    if (vs) {
                        for (int w = lane; w < W; w += 32) cp_async_4(stage + s * W + w, gst + gs * W + w);
                    }
                }
            }
            cp_async_commit();
        };

        // the first chunk is requested BEFORE the outputs of the games at rest are written: pass 2 runs in the shadow of
        // that load's DRAM latency (late in a rollout a warp has one short chunk: the load was the longest wait of the step)
        long long g_cur = 0, g_nxt = 0;
        bool v_cur = false, v_nxt = false;
        if (spec && !natural) {  // the guess was wrong: let the tile land, then forget it
            mbar_wait(&bar[0], phase_bits & 1u);
            phase_bits ^= 1u;
            spec = false;
        }
        if (inplay) {
            chunk_game(0, g_cur, v_cur);
            if (!spec) gather(g_cur, v_cur, stages, 0);  // (natural order and a good guess: chunk 0 IS the tile in flight)
        }
        // pass 2 (fused observation only): the outputs of the games at rest, from their census bytes
        for (int k = 0; OBS && k < nk; ++k) {
            const long long g = (tile_of(j0 + k) << 5) + lane;
            const uint32_t code = clsb[k * 32 + lane];
            const bool rest = code & 0x80u;
            const uint32_t v = (code >> 6) & 1u;  // live count of a game at rest
            if (rest) {
                if (p.done) p.done[g] = 1;
                if (p.reward) p.reward[g] = rest_reward;
                if (p.num_points) p.num_points[g] = (int32_t)v;
            }
            const uint32_t restmask = __ballot_sync(0xffffffffu, rest);
            settled_total += __popc(restmask);
            if constexpr (OBS) {
                // (warp-uniform) the constant observation of every game at rest of this tile — unless the tile is about to
                // be stepped whole in natural order: its observation rows are then stored as one block, these included
                if (p.obs && restmask && !(natural && ((tilemask >> k) & 1u))) {
                    // staged in the observation tile like a chunk's rows, then ONE bulk store for a tile that is
                    // at rest as a whole (the common case late in a rollout), one per game otherwise
                    float* orow = obs_tile + lane * OW;
                    bulk_wait_read<0>();  // earlier observation stores have read the tile
                    __syncwarp();
                    if (rest) {
                        const float z = (v & 1u) ? 0.0f : p.pad;  // the lone point, at the origin, sorts first
                        const uint32_t ocm = p.obs_coord ? action_mask(load_action(p.obs_coord, g, p.flags), p.flags) : 0u;
                        if ((OW & 3) == 0 && D <= 4 && !p.obs_coord) {
                            float4* o4 = reinterpret_cast<float4*>(orow);
                            const float4 padv4 = make_float4(p.pad, p.pad, p.pad, p.pad);
                            o4[0] = make_float4(z, D > 1 ? z : p.pad, D > 2 ? z : p.pad, D > 3 ? z : p.pad);
#pragma unroll
                            for (int q = 1; q < W / 4; ++q) o4[q] = padv4;
                        } else {
                            for (int w = 0; w < OW; ++w)
                                orow[w] = (w < D) ? z : (w < W ? p.pad : (float)((ocm >> (w - W)) & 1u));
                        }
                    }
                    const long long first = g - lane;
                    const int cnt = (int)((B - first < 32) ? (B - first) : 32);
                    const uint32_t allmask = (cnt == 32) ? 0xffffffffu : ((1u << cnt) - 1u);
                    if (obs_bulk) {
                        fence_async_smem();
                        __syncwarp();
                        if (restmask == allmask) {
                            if (lane == 0) bulk_store(p.obs + first * OW, obs_tile, (uint32_t)(cnt * OW) * 4u);
                        } else if (rest) {
                            bulk_store(p.obs + g * OW, orow, (uint32_t)OW * 4u);
                        }
                        bulk_commit();
                    } else {
                        __syncwarp();
                        uint32_t m = restmask;
                        while (m) {
                            const int r = __ffs((int)m) - 1;
                            m &= m - 1;
                            warp_copy_words(reinterpret_cast<uint32_t*>(p.obs + (first + r) * OW),
                                            reinterpret_cast<uint32_t*>(obs_tile + r * OW), OW, lane);
                        }
                    }
                    __syncwarp();
                }
            }
            // the tile's word of the done mask: the games at rest now, the games that finish in this step later
            if (p.done_bits && lane == 0) p.done_bits[tile_of(j0 + k)] = restmask;
        }
        __syncwarp();
        if (inplay == 0) continue;
        for (int v = 0; v < nchunks; ++v) {
            const int sidx = (STAGES == 1) ? 0 : (v & 1);
            uint32_t* stage = stages + sidx * L::STAGE_WORDS;
            LaneState ls;
            ls.g = g_cur;
            ls.valid = v_cur;
            ls.shift = (p.ops & HK_OP_SHIFT) && ls.valid;
            ls.ha = 3;
            ls.ax = 0;
            ls.origin = false;
            if (ls.shift) load_actions(p, p.flags, ls.g, 0, ls.ha, ls.ax);
            if constexpr (STAGES == 1 && HK_L2_PREFETCH) {
                // one stage: the next tile cannot be requested before this one has been stored, but it can wait in L2
                if (bulk && natural && tiles_left && lane == 0) {
                    const long long tn = tile_of(j0 + (__ffs((int)tiles_left) - 1));
                    const long long left = B - (tn << 5);
                    bulk_prefetch_l2(gst + (tn << 5) * W, (uint32_t)(left < 32 ? left : 32) * (uint32_t)(W * 4));
                }
            }
            if (bulk) {  // this chunk's games have landed (they were requested one chunk ago)
                mbar_wait(&bar[sidx], (phase_bits >> sidx) & 1u);
                phase_bits ^= (1u << sidx);
            }
            if constexpr (STAGES >= 2) {
                // the next chunk streams into the other stage while this one is processed (its bulk stores of
                // the chunk before have had the wait above to drain)
                if (v + 1 < nchunks) {
                    chunk_game(v + 1, g_nxt, v_nxt);
                    gather(g_nxt, v_nxt, stages + ((v + 1) & 1) * L::STAGE_WORDS, (v + 1) & 1);
                    if (!bulk) cp_async_wait<1>();
                } else {
                    if (!bulk) cp_async_wait<0>();
                }
            } else {
                if (!bulk) cp_async_wait<0>();
            }
            __syncwarp();
            uint32_t* row = stage + lane * W;
            bool exceed = false;
            bool chg = false;  // in place: only changed games go back
            bool released = false;  // the stage has been handed back (and refilled) in the middle of the step
            if constexpr (SCHED_DIRECT && !OBS && STAGES == 1 && !Elem<T>::is_float) {
                // packed tiers, direct route (hk_small.cuh): once the rows are gathered the stage is refilled — the next
                // chunk's load runs behind this chunk's arithmetic — and the changed rows go straight to global memory
                auto release = [&]() {
                    released = true;
                    __syncwarp();
                    if (v + 1 < nchunks) {
                        chunk_game(v + 1, g_nxt, v_nxt);
                        gather(g_nxt, v_nxt, stages, 0);
                    }
                };
                small_process_tile<T, N, D, false, true, decltype(release)>(p, ls, row, exceed, chg, gst + ls.g * W,
                                                                            release);
            } else {
                small_process_tile<T, N, D, false>(p, ls, row, exceed, chg);
            }
            if (ls.valid) {
                if (p.num_points) p.num_points[ls.g] = ls.cnt;
                if (p.done_bits && ls.cnt < 2) atomicOr(p.done_bits + (ls.g >> 5), 1u << (ls.g & 31));
                const uint32_t nv = (ls.cnt <= 1) ? (0x80u | (ls.origin ? 2u : 0u) | (uint32_t)ls.cnt)
                                                  : (uint32_t)(ls.cnt > 127 ? 127 : ls.cnt);
                p.census[ls.g] = (uint8_t)nv;
            }
            if (p.exceed_flag) {
                if (__any_sync(0xffffffffu, exceed && ls.valid) && lane == 0) *p.exceed_flag = 1;
            }
            if constexpr (OBS) {
                if (p.obs) {  // ---- observation features of the new state (in the lane's row area), as in K-small ----
                    const bool ingrid = ls.g < B;  // (natural order: the games at rest of the tile ride along)
                    float* orow = obs_tile + lane * OW;
                    bulk_wait_read<0>();  // the previous chunk's observation stores have read the tile
                    __syncwarp();
                    {
                        T z[W];
                        load_game<T, W>(row, z);
                        if constexpr (Elem<T>::is_float) {
#pragma unroll
                            for (int q = 0; q < W; ++q) z[q] = z[q] + 0.0f;
                        }
                        const uint32_t zl = live_mask<T, N, D>(z);
                        const int zmax = __reduce_max_sync(0xffffffffu, ingrid ? __popc(zl) : 0);
                        bool built = false;
                        if constexpr (D <= 6) built = features_network<T, N, D>(z, zl, zmax, p.flags, p.pad, orow);
                        if (!built) {
                            const int gpow = (OW & -OW) > 32 ? 32 : (OW & -OW);
                            features_rolled<T, N, D>(row, zl, zmax, p.flags, p.pad, orow, (lane * gpow) >> 5);
                        }
                    }
                    if (p.obs_coord) {
                        const uint32_t ocm = ingrid ? action_mask(load_action(p.obs_coord, ls.g, p.flags), p.flags) : 0u;
#pragma unroll
                        for (int k = 0; k < D; ++k) orow[W + k] = (float)((ocm >> k) & 1u);
                    }
                    if (natural) {  // the tile's observations are contiguous
                        const long long first = ls.g - lane;
                        const int cnt = (int)((B - first < 32) ? (B - first) : 32);
                        if (obs_bulk) {
                            fence_async_smem();
                            __syncwarp();
                            if (lane == 0) bulk_store(p.obs + first * OW, obs_tile, (uint32_t)(cnt * OW) * 4u);
                            bulk_commit();
                        } else {
                            __syncwarp();
                            warp_copy_words(reinterpret_cast<uint32_t*>(p.obs + first * OW), reinterpret_cast<uint32_t*>(obs_tile),
                                            cnt * OW, lane);
                        }
                    } else if (obs_bulk) {  // one bulk store per game
                        fence_async_smem();
                        __syncwarp();
                        if (ls.valid) bulk_store(p.obs + ls.g * OW, orow, (uint32_t)OW * 4u);
                        bulk_commit();
                    } else {
                        __syncwarp();
                        for (int s2 = 0; s2 < 32; ++s2) {
                            const long long gs = __shfl_sync(0xffffffffu, ls.g, s2);
                            const bool vs = __shfl_sync(0xffffffffu, ls.valid ? 1 : 0, s2) != 0;
                            if (vs) warp_copy_words(reinterpret_cast<uint32_t*>(p.obs + gs * OW),
                                                    reinterpret_cast<uint32_t*>(obs_tile + s2 * OW), OW, lane);
                        }
                    }
                    __syncwarp();
                }
            }
            // ---- write-back of the changed games, GPI games per instruction ----
            uint32_t dirty = __ballot_sync(0xffffffffu, ls.valid && chg);
            __syncwarp();
            if (natural && __popc(dirty) > HK_SPARSE_STORE_MAX) {  // many changed games: the tile goes back whole
                const long long first = ls.g - lane;
                const int cnt = (int)((B - first < 32) ? (B - first) : 32);
                if (bulk) {
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) bulk_store(gst + first * W, stage, (uint32_t)cnt * (uint32_t)(W * 4));
                    bulk_commit();
                } else {
                    warp_copy_words(gst + first * W, stage, cnt * W, lane);
                }
            } else if (bulk) {
                if (dirty) {
                    fence_async_smem();  // the lanes' st.shared results before the async-proxy reads of the stores
                    __syncwarp();
                    if (ls.valid && chg) bulk_store(gst + ls.g * W, row, (uint32_t)(W * 4));
                    bulk_commit();
                }
            } else if (piece_ok) {
                while (dirty) {
                    int s = -1;
                    uint32_t m = dirty;
#pragma unroll
                    for (int q = 0; q < GPI; ++q) {  // the q-th dirty game goes to the lanes with sub == q
                        const int b = m ? (__ffs((int)m) - 1) : -1;
                        m &= m - 1;
                        if (q == sub) s = b;
                    }
                    dirty = m;
                    const long long gs = __shfl_sync(0xffffffffu, ls.g, s < 0 ? 0 : s);
                    if (sub < GPI && s >= 0) {
                        const uint32_t* src = stage + s * W + piece * CHW;
                        uint32_t* dst = gst + gs * W + piece * CHW;
                        if constexpr (CHW == 4) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
                        else if constexpr (CHW == 2) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(src);
                        else *dst = *src;
                    }
                }
            } else {
                while (dirty) {
                    const int s = __ffs((int)dirty) - 1;
                    dirty &= dirty - 1;
                    const long long gs = __shfl_sync(0xffffffffu, ls.g, s);
                    warp_copy_words(gst + gs * W, stage + s * W, W, lane);
                }
            }
            __syncwarp();  // every lane is done with this stage before a later gather overwrites it
            if constexpr (STAGES == 1) {
                if (v + 1 < nchunks && !released) {
                    chunk_game(v + 1, g_nxt, v_nxt);
                    gather(g_nxt, v_nxt, stages, 0);
                }
            }
            g_cur = g_nxt;
            v_cur = v_nxt;
        }
    }
    if (bulk || obs_bulk) bulk_wait_all<0>();  // this lane's bulk stores are globally complete before the warp retires
    if (p.done_count && settled_total && lane == 0) atomicAdd(p.done_count, settled_total);
}

}  // namespace hk
