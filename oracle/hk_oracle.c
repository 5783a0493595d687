/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Plain-C restatement of the reference env step.
 *
 * Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg load this library (through oracle/cport.py).  It exists so that parity
 * can be checked at BASELINE sizes (1 M games) where the NumPy restatement's [B,N,N,d]
 * temporaries do not fit, and so that a CPU number can be timed on the GPU box's host cores.
 *
 * Each block cites the reference function it restates (paths relative to the reference root).
 * The loop structure is the reference's algorithm (all-pairs O(N^2 d) dedupe + dominance,
 * per-game independent), written the way hironaka/cpp/cppUtil.cpp:4-56 writes its own
 * brute-force loop; games are distributed over pthreads (no libgomp in this image).
 *
 * Pinning: tests/test_oracle_golden.py checks this port against the reference's known-answer
 * vectors (test/testTensorPoints.py, test/testJAX.py) and against outputs of the real
 * reference (tests/golden/ref_*.npz), and tests/test_oracle_cross.py against hk_oracle.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#include "../include/hironaka_b200.h"

#define ORACLE_MAX_WORDS HK_MAX_GAME_WORDS

/* host_action -> coordinate bitmask.  Discrete id k = the k-th integer >= 3 that is not a
 * power of two (HostActionEncoder.__init__, hironaka/src/_fn.py:255-269;
 * decode_table, hironaka/jax/host_action_preprocess.py:8-24). */
static uint32_t oracle_decode(int32_t id) {
    int32_t k = -1;
    for (uint32_t m = 1;; ++m) {
        if ((m & (m - 1)) == 0) continue;
        if (++k == id) return m;
    }
}

/* ---- a minimal pthread parallel-for over games (this image has no libgomp) ------------- */
static int g_threads = 0;

int hk_oracle_threads(void) {
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

void hk_oracle_set_threads(int n) { g_threads = n > 0 ? n : 0; }

typedef void (*range_fn)(int64_t lo, int64_t hi, void* ctx);
typedef struct {
    range_fn fn;
    void* ctx;
    int64_t lo, hi;
} range_job;

static void* range_thread(void* p) {
    range_job* j = (range_job*)p;
    j->fn(j->lo, j->hi, j->ctx);
    return 0;
}

static void parallel_for(int64_t B, range_fn fn, void* ctx) {
    int T = hk_oracle_threads();
    if (T > 256) T = 256;
    if (B < 4096 || T <= 1) {
        fn(0, B, ctx);
        return;
    }
    pthread_t th[256];
    range_job jobs[256];
    int64_t chunk = (B + T - 1) / T;
    int started = 0;
    for (int t = 0; t < T; ++t) {
        int64_t lo = t * chunk, hi = lo + chunk > B ? B : lo + chunk;
        if (lo >= hi) break;
        jobs[t].fn = fn;
        jobs[t].ctx = ctx;
        jobs[t].lo = lo;
        jobs[t].hi = hi;
        if (pthread_create(&th[t], 0, range_thread, &jobs[t]) != 0) {
            fn(lo, hi, ctx); /* fall back to running this chunk inline */
            th[t] = 0;
        }
        started = t + 1;
    }
    for (int t = 0; t < started; ++t)
        if (th[t]) pthread_join(th[t], 0);
}

#include <math.h>
/* Zeillinger's host (zeillinger_fn_slice hironaka/jax/players.py:55-105) on one game, float32 */
static uint32_t oracle_zeillinger(const float* pts, int N, int d) {
    float bestL = INFINITY, bestS = INFINITY;
    int bi = -1, bj = -1;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            int neg = 0;
            for (int k = 0; k < d; ++k) neg |= (pts[i * d + k] < 0) || (pts[j * d + k] < 0);
            if (neg) continue;
            float mx = pts[i * d] - pts[j * d], mn = mx;
            for (int k = 1; k < d; ++k) {
                float df = pts[i * d + k] - pts[j * d + k];
                if (df > mx) mx = df;
                if (df < mn) mn = df;
            }
            if (fabsf(mx - mn) <= 1e-8f + 1e-5f * fabsf(mn)) continue; /* jnp.isclose(maximal, minimal) */
            int cmax = 0, cmin = 0;
            for (int k = 0; k < d; ++k) {
                float df = pts[i * d + k] - pts[j * d + k];
                cmax += df == mx;
                cmin += df == mn;
            }
            float L = mx - mn, S = (float)(mx == mn ? cmax : cmax + cmin);
            if (L < bestL || (L == bestL && S < bestS)) { /* lexsort is stable: first in flat order wins ties */
                bestL = L;
                bestS = S;
                bi = i;
                bj = j;
            }
        }
    if (bi < 0) return 3u; /* every pair is (inf, inf): index 0 = pair (0,0), constant difference -> action 0 = {0,1} */
    int amin = 0, amax = 0;
    float mn = pts[bi * d] - pts[bj * d], mx = mn;
    for (int k = 1; k < d; ++k) {
        float df = pts[bi * d + k] - pts[bj * d + k];
        if (df < mn) { mn = df; amin = k; }
        if (df > mx) { mx = df; amax = k; }
    }
    return amin == amax ? 3u : ((1u << amin) | (1u << amax));
}

#define NO_RESCALE(x, N, d, pad) (void)0
#define F32_RESCALE(x, N, d, pad)                                                                       \
    do { /* rescale_torch hironaka/src/_torch_ops.py:136-146; with HK_F_RESCALE_EPS the rule of     \
            calculate_rescale hironaka/src/_jax_ops.py:93-98: a maximum <= 1e-8 leaves the game alone */ \
        float mx_ = x[0];                                                                               \
        for (int t = 1; t < N * d; ++t) mx_ = x[t] > mx_ ? x[t] : mx_;                                  \
        if (mx_ == 0.0f || ((flags & HK_F_RESCALE_EPS) && mx_ <= 1e-8f)) mx_ = 1.0f;                    \
        for (int i = 0; i < N; ++i) {                                                                   \
            int lv_ = x[i * d] >= 0;                                                                    \
            for (int k = 0; k < d; ++k) x[i * d + k] = lv_ ? x[i * d + k] / mx_ : pad;                  \
        }                                                                                               \
    } while (0)

#define DEFINE_STEP(NAME, T, RESC, PADV)                                                                \
    static void NAME##_one(const T* in, T* x, uint32_t cmask, int32_t a, int N, int d, uint32_t ops,    \
                           uint32_t flags, T pad, uint8_t* done, float* reward, int32_t* num_points) { \
        int live_before = 0;                                                                            \
        for (int i = 0; i < N; ++i) {                                                                   \
            live_before += in[i * d] >= 0;                                                              \
            for (int k = 0; k < d; ++k) x[i * d + k] = in[i * d + k];                                   \
        }                                                                                               \
        int prev_done = live_before < 2; /* get_dones, hironaka/jax/util.py:34-35 */                   \
        if (ops & HK_OP_SHIFT) {                                                                        \
            /* fixed players: all_coord_host_fn / zeillinger_fn / choose_first|last_agent_fn            \
               (hironaka/jax/players.py:42-52,55-105,156-212) */                                        \
            if (flags & HK_F_HOST_ALL_COORD) cmask = (1u << d) - 1u;                                    \
            else if (flags & HK_F_HOST_ZEILLINGER) {                                                    \
                float zf[ORACLE_MAX_WORDS];                                                             \
                for (int t = 0; t < N * d; ++t) zf[t] = (float)x[t];                                    \
                cmask = oracle_zeillinger(zf, N, d);                                                    \
            }                                                                                           \
            if (flags & HK_F_AGENT_FIRST) { a = 0; while (a < d - 1 && !((cmask >> a) & 1u)) ++a; if (!cmask) a = 0; } \
            else if (flags & HK_F_AGENT_LAST) { a = d - 1; while (a > 0 && !((cmask >> a) & 1u)) --a; if (!cmask) a = d - 1; } \
            /* shift_torch hironaka/src/_torch_ops.py:46-110; shift_jax _jax_ops.py:76-90 */           \
            int apply = 1;                                                                              \
            if ((flags & HK_F_NOOP_INVALID) && !((cmask >> a) & 1u)) apply = 0;   /* :90-91 */          \
            if ((flags & HK_F_FREEZE_ENDED) && prev_done) apply = 0;              /* :92-93 */          \
            if (a < 0 || a >= d) apply = 0; /* arange(d)==axis never matches, _jax_ops.py:79 */         \
            for (int i = 0; i < N; ++i) {                                                               \
                if (x[i * d] >= 0) {                                                                    \
                    if (apply) {                                                                        \
                        T s = 0;                                                                        \
                        for (int k = 0; k < d; ++k)                                                     \
                            if ((cmask >> k) & 1u) s = s + x[i * d + k];                                \
                        x[i * d + a] = s;                                                               \
                    }                                                                                   \
                } else {                                                                                \
                    for (int k = 0; k < d; ++k) x[i * d + k] = pad;               /* :104 */            \
                }                                                                                       \
            }                                                                                           \
        }                                                                                               \
        if (ops & HK_OP_REPOSITION) {                                                                   \
            /* reposition_torch _torch_ops.py:113-133; subtract_min _jax_ops.py:114-120 */             \
            for (int k = 0; k < d; ++k) {                                                               \
                int any = 0;                                                                            \
                T mn = 0;                                                                               \
                for (int i = 0; i < N; ++i)                                                             \
                    if (x[i * d] >= 0) {                                                                \
                        if (!any || x[i * d + k] < mn) mn = x[i * d + k];                               \
                        any = 1;                                                                        \
                    }                                                                                   \
                for (int i = 0; i < N; ++i) {                                                           \
                    if (x[i * d] >= 0) x[i * d + k] = x[i * d + k] - mn;                                \
                }                                                                                       \
            }                                                                                           \
            for (int i = 0; i < N; ++i)                                                                 \
                if (x[i * d] < 0)                                                                       \
                    for (int k = 0; k < d; ++k) x[i * d + k] = pad;                                     \
        }                                                                                               \
        if (ops & HK_OP_DEDUPE) { /* remove_repeated hironaka/src/_fn.py:192-213 */                     \
            uint8_t rep[HK_MAX_POINTS];                                                                 \
            for (int i = 0; i < N; ++i) {                                                               \
                rep[i] = 0;                                                                             \
                for (int j = 0; j < i && !rep[i]; ++j) {                                                \
                    int eq = 1;                                                                         \
                    for (int k = 0; k < d; ++k)                                                         \
                        if (x[i * d + k] != x[j * d + k]) {                                             \
                            eq = 0;                                                                     \
                            break;                                                                      \
                        }                                                                               \
                    rep[i] = (uint8_t)eq;                                                               \
                }                                                                                       \
            }                                                                                           \
            for (int i = 0; i < N; ++i)                                                                 \
                if (rep[i] || x[i * d] < 0)                                                             \
                    for (int k = 0; k < d; ++k) x[i * d + k] = pad;                                     \
        }                                                                                               \
        if (ops & HK_OP_NEWTON) {                                                                       \
            /* remove_repeated hironaka/src/_fn.py:192-213, then                                        \
               get_newton_polytope_approx_torch _torch_ops.py:8-39 (get_interior _jax_ops.py:43-57):    \
               both passes read the state BEFORE their own removals. */                                 \
            uint8_t rep[HK_MAX_POINTS];                                                                 \
            uint8_t rem[HK_MAX_POINTS];                                                                 \
            for (int i = 0; i < N; ++i) {                                                               \
                rep[i] = 0;                                                                             \
                for (int j = 0; j < i && !rep[i]; ++j) {                                                \
                    int eq = 1;                                                                         \
                    for (int k = 0; k < d; ++k)                                                         \
                        if (x[i * d + k] != x[j * d + k]) {                                             \
                            eq = 0;                                                                     \
                            break;                                                                      \
                        }                                                                               \
                    rep[i] = (uint8_t)eq;                                                               \
                }                                                                                       \
            }                                                                                           \
            for (int i = 0; i < N; ++i)                                                                 \
                if (rep[i])                                                                             \
                    for (int k = 0; k < d; ++k) x[i * d + k] = pad;                                     \
            for (int i = 0; i < N; ++i) {                                                               \
                rem[i] = 0;                                                                             \
                if (x[i * d] < 0) continue;                                                             \
                for (int j = 0; j < N && !rem[i]; ++j) {                                                \
                    if (j == i || x[j * d] < 0) continue;                                               \
                    int ge = 1;                                                                         \
                    for (int k = 0; k < d; ++k)                                                         \
                        if (x[i * d + k] - x[j * d + k] < 0) {                                          \
                            ge = 0;                                                                     \
                            break;                                                                      \
                        }                                                                               \
                    rem[i] = (uint8_t)ge;                                                               \
                }                                                                                       \
            }                                                                                           \
            for (int i = 0; i < N; ++i)                                                                 \
                if (rem[i] || x[i * d] < 0)                                                             \
                    for (int k = 0; k < d; ++k) x[i * d + k] = pad;                                     \
        }                                                                                               \
        if (ops & HK_OP_RESCALE) RESC(x, N, d, pad);                                                    \
        int live_after = 0;                                                                             \
        for (int i = 0; i < N; ++i) live_after += x[i * d] >= 0;                                        \
        int dn = live_after < 2;                                                                        \
        if (done) *done = (uint8_t)dn;                                                                  \
        if (reward) { /* get_reward_fn hironaka/jax/util.py:128-149 */                                  \
            float r = (dn && !prev_done) ? 1.0f : 0.0f;                                                 \
            *reward = (flags & HK_F_ROLE_AGENT) ? -r : r;                                               \
        }                                                                                               \
        if (num_points) *num_points = live_after;                                                       \
    }                                                                                                   \
                                                                                                        \
    typedef struct {                                                                                    \
        const T* state_in; T* state_out; const int32_t* host_action; const int32_t* axis;               \
        uint8_t* done; float* reward; int32_t* num_points; int32_t N, d; uint32_t ops, flags; T pad;    \
    } NAME##_ctx;                                                                                       \
    static void NAME##_range(int64_t lo, int64_t hi, void* vp) {                                        \
        NAME##_ctx* c = (NAME##_ctx*)vp;                                                                \
        const int N = c->N, d = c->d;                                                                   \
        T* x = (T*)malloc(sizeof(T) * (size_t)N * (size_t)d);                                           \
        for (int64_t b = lo; b < hi; ++b) {                                                             \
            uint32_t cm = 0;                                                                            \
            int32_t a = 0;                                                                              \
            if (c->ops & HK_OP_SHIFT) {                                                                 \
                if (c->host_action)                                                                     \
                    cm = (c->flags & HK_F_ACT_DISCRETE) ? oracle_decode(c->host_action[b])              \
                                                        : (uint32_t)c->host_action[b];                  \
                if (c->axis) a = c->axis[b];                                                            \
            }                                                                                           \
            NAME##_one(c->state_in + b * N * d, x, cm, a, N, d, c->ops, c->flags, c->pad,               \
                       c->done ? c->done + b : 0, c->reward ? c->reward + b : 0,                        \
                       c->num_points ? c->num_points + b : 0);                                          \
            memcpy(c->state_out + b * N * d, x, sizeof(T) * (size_t)N * (size_t)d);                     \
        }                                                                                               \
        free(x);                                                                                        \
    }                                                                                                   \
    int NAME(const T* state_in, T* state_out, const int32_t* host_action, const int32_t* axis,          \
             uint8_t* done, float* reward, int32_t* num_points, int64_t B, int32_t N, int32_t d,        \
             uint32_t ops, uint32_t flags, float padding_value) {                                       \
        if (N < 1 || N > HK_MAX_POINTS || d < 1 || d > 31 || (int64_t)N * d > ORACLE_MAX_WORDS)         \
            return HK_ERR_UNSUPPORTED;                                                                  \
        NAME##_ctx c = {state_in, state_out, host_action, axis, done, reward, num_points,               \
                        N, d, ops, flags, PADV};                                                        \
        parallel_for(B, NAME##_range, &c);                                                              \
        return HK_OK;                                                                                   \
    }

DEFINE_STEP(hk_oracle_step_i32, int32_t, NO_RESCALE, (int32_t)padding_value)
DEFINE_STEP(hk_oracle_step_f32, float, F32_RESCALE, padding_value)

/* rescale_torch hironaka/src/_torch_ops.py:136-146 (== rescale_jax _jax_ops.py:93-111 on
 * well-formed input): live entries divided by the game max (0 -> 1) in IEEE float32. */
int hk_oracle_rescale_f32(const float* in, float* out, int64_t B, int32_t N, int32_t d, float pad) {
    for (int64_t b = 0; b < B; ++b) {
        const float* x = in + b * N * d;
        float* y = out + b * N * d;
        float mx = x[0];
        for (int t = 1; t < N * d; ++t) mx = x[t] > mx ? x[t] : mx;
        if (mx == 0.0f) mx = 1.0f;
        for (int i = 0; i < N; ++i)
            for (int k = 0; k < d; ++k) y[i * d + k] = x[i * d] >= 0 ? x[i * d + k] / mx : pad;
    }
    return HK_OK;
}

/* Observation features of an [B,N,d] state, from either dtype, as float32 [B, N*d (+d)]:
 * optional rescale (util.py:183 / _torch_ops.py:136-146), then a STABLE descending row sort
 * by coordinate 0 (tensor_points.py:72-74) or lexicographic with the last coordinate primary
 * (util.py:195), then optional coordinate set appended (make_agent_obs util.py:22-31). */
static int lex_before(const float* a, int ia, const float* b, int ib, int d, int lex) {
    /* 1 if row a sorts strictly before row b (descending keys, stable on index);
       lex: 1 = last coordinate primary (util.py:195), 2 = coordinate 0 primary (_list_ops.py:25) */
    if (lex == 1) {
        for (int k = d - 1; k >= 0; --k) {
            if (a[k] > b[k]) return 1;
            if (a[k] < b[k]) return 0;
        }
    } else if (lex == 2) {
        for (int k = 0; k < d; ++k) {
            if (a[k] > b[k]) return 1;
            if (a[k] < b[k]) return 0;
        }
    } else {
        if (a[0] > b[0]) return 1;
        if (a[0] < b[0]) return 0;
    }
    return ia < ib;
}

static void features_one(const float* f, float* o, int N, int d, uint32_t flags) {
    int sorted = (flags & (HK_F_OBS_SORT_COORD0 | HK_F_OBS_SORT_LEX | HK_F_OBS_SORT_LEX_FIRST)) != 0;
    int lex = (flags & HK_F_OBS_SORT_LEX) ? 1 : ((flags & HK_F_OBS_SORT_LEX_FIRST) ? 2 : 0);
    for (int i = 0; i < N; ++i) {
        int r = i;
        if (sorted) {
            r = 0;
            for (int j = 0; j < N; ++j)
                if (j != i && lex_before(f + j * d, j, f + i * d, i, d, lex)) ++r;
        }
        for (int k = 0; k < d; ++k) o[r * d + k] = f[i * d + k];
    }
}

#define DEFINE_FEATURES(NAME, T)                                                                       \
    typedef struct {                                                                                   \
        const T* state; float* obs; const int32_t* obs_coord; int32_t N, d; uint32_t flags; float pad; \
    } NAME##_ctx;                                                                                      \
    static void NAME##_range(int64_t lo, int64_t hi, void* vp) {                                       \
        NAME##_ctx* c = (NAME##_ctx*)vp;                                                               \
        const int N = c->N, d = c->d;                                                                  \
        const uint32_t flags = c->flags;                                                               \
        const int W = N * d + (c->obs_coord ? d : 0);                                                  \
        float* f = (float*)malloc(sizeof(float) * (size_t)N * (size_t)d);                              \
        for (int64_t b = lo; b < hi; ++b) {                                                            \
            const T* x = c->state + b * N * d;                                                         \
            float mx = (float)x[0];                                                                    \
            for (int t = 1; t < N * d; ++t) mx = (float)x[t] > mx ? (float)x[t] : mx;                  \
            if (mx == 0.0f || ((flags & HK_F_RESCALE_EPS) && mx <= 1e-8f)) mx = 1.0f;                  \
            for (int i = 0; i < N; ++i)                                                                \
                for (int k = 0; k < d; ++k) {                                                          \
                    float v = (float)x[i * d + k];                                                     \
                    if (x[i * d] < 0) v = c->pad;                                                      \
                    else if (flags & HK_F_OBS_RESCALE) v = v / mx;                                     \
                    f[i * d + k] = v;                                                                  \
                }                                                                                      \
            features_one(f, c->obs + b * W, N, d, flags);                                              \
            if (c->obs_coord) {                                                                        \
                uint32_t cm = (flags & HK_F_ACT_DISCRETE) ? oracle_decode(c->obs_coord[b])             \
                                                          : (uint32_t)c->obs_coord[b];                 \
                for (int k = 0; k < d; ++k) c->obs[b * W + N * d + k] = (float)((cm >> k) & 1u);       \
            }                                                                                          \
        }                                                                                              \
        free(f);                                                                                       \
    }                                                                                                  \
    int NAME(const T* state, float* obs, const int32_t* obs_coord, int64_t B, int32_t N, int32_t d,    \
             uint32_t flags, float pad) {                                                              \
        if ((int64_t)N * d > ORACLE_MAX_WORDS) return HK_ERR_UNSUPPORTED;                              \
        NAME##_ctx c = {state, obs, obs_coord, N, d, flags, pad};                                      \
        parallel_for(B, NAME##_range, &c);                                                             \
        return HK_OK;                                                                                  \
    }

DEFINE_FEATURES(hk_oracle_features_i32, int32_t)
DEFINE_FEATURES(hk_oracle_features_f32, float)
