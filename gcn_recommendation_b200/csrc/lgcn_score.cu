// lgcn_score.cu -- full-rank rating: scores + train-item mask + per-user top-k (sm_100a).
//
// Replaces torch.matmul(user_batch, item_table.T) (reference main.py:420), the per-user
// python mask loop (reference main.py:422-424) and torch.topk (reference main.py:426).
// The [users x items] score matrix never reaches HBM: a CTA owns a tile of 64 users, sweeps
// the item table in tiles of 128, and keeps every user's running top-k in shared memory.
//
// This is the EXACT fp32 path: score = sequential fp32 FMA over the feature index (the order
// of the CPU oracle), so ids are bit-reproducible.  (The tcgen05 filter pass that feeds this
// exact re-score for the large-catalogue sweep lives in lgcn_score_tc.cu.)
#include <float.h>

#include "lgcn_common.cuh"

namespace lgcn {

constexpr int TU = 64;     // users per CTA
constexpr int TI = 128;    // items per tile
constexpr int KC = 64;     // feature chunk staged in smem
constexpr int KMAX = 32;   // top-k capacity (one warp lane per entry)
constexpr int kScoreThreads = 256;
constexpr int AS_LD = TU + 4;
constexpr int BS_LD = TI + 4;

struct ScoreSmem {
    float As[KC][AS_LD];
    float Bs[KC][BS_LD];
    float q_s[TU][TI];
    int q_i[TU][TI];
    float tk_s[TU][KMAX];
    int tk_i[TU][KMAX];
    float thr[TU];
    int q_cnt[TU];
    long long urow[TU];
};

__device__ __forceinline__ bool better(float s, int id, float ts, int tid) {
    return s > ts || (s == ts && id < tid);
}

__global__ void __launch_bounds__(kScoreThreads, 1)
score_topk_kernel(const float *__restrict__ Fu, const float *__restrict__ Fi,
                  const int64_t *__restrict__ users, int64_t nu, int64_t n_items, int d,
                  const int64_t *__restrict__ mask_rowptr, const int32_t *__restrict__ mask_col,
                  int k, int32_t *__restrict__ out_ids, float *__restrict__ out_scores,
                  int64_t items_per_split) {
    // blockIdx.y = item split: this CTA sweeps items [y*items_per_split, (y+1)*items_per_split)
    // and writes its partial top-k at split offset y (merged by merge_topk_kernel).
    const int64_t it_begin = (int64_t)blockIdx.y * items_per_split;
    const int64_t it_end = min(n_items, it_begin + items_per_split);
    out_ids += (size_t)blockIdx.y * nu * k;
    out_scores += (size_t)blockIdx.y * nu * k;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScoreSmem &sm = *reinterpret_cast<ScoreSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.x * TU;
    const int tx = tid & 31;   // items tx*4 .. tx*4+3
    const int ty = tid >> 5;   // users ty*8 .. ty*8+7

    for (int i = tid; i < TU; i += kScoreThreads) {
        const int64_t q = q0 + i;
        sm.urow[i] = (q < nu) ? (long long)users[q] : -1;
        sm.thr[i] = -FLT_MAX;
        sm.q_cnt[i] = 0;
    }
    for (int i = tid; i < TU * KMAX; i += kScoreThreads) {
        sm.tk_s[i / KMAX][i % KMAX] = -FLT_MAX;
        sm.tk_i[i / KMAX][i % KMAX] = -1;
    }
    __syncthreads();

    for (int64_t it0 = it_begin; it0 < it_end; it0 += TI) {
        float acc[8][4];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

        for (int k0 = 0; k0 < d; k0 += KC) {
            const int kc = min(KC, d - k0);
            // stage the user chunk transposed: As[j][u]
            for (int idx = tid; idx < TU * (KC / 4); idx += kScoreThreads) {
                const int u = idx % TU, jq = idx / TU;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const long long r = sm.urow[u];
                if (r >= 0 && jq * 4 < kc) v = ld_nc_f4(Fu + (size_t)r * d + k0 + jq * 4);
                sm.As[jq * 4 + 0][u] = v.x;
                sm.As[jq * 4 + 1][u] = v.y;
                sm.As[jq * 4 + 2][u] = v.z;
                sm.As[jq * 4 + 3][u] = v.w;
            }
            // stage the item chunk transposed: Bs[j][i]
            for (int idx = tid; idx < TI * (KC / 4); idx += kScoreThreads) {
                const int i = idx % TI, jq = idx / TI;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int64_t item = it0 + i;
                if (item < it_end && jq * 4 < kc) v = ld_nc_f4(Fi + (size_t)item * d + k0 + jq * 4);
                sm.Bs[jq * 4 + 0][i] = v.x;
                sm.Bs[jq * 4 + 1][i] = v.y;
                sm.Bs[jq * 4 + 2][i] = v.z;
                sm.Bs[jq * 4 + 3][i] = v.w;
            }
            __syncthreads();
#pragma unroll 4
            for (int j = 0; j < kc; ++j) {
                const float4 a0 = *reinterpret_cast<const float4 *>(&sm.As[j][ty * 8]);
                const float4 a1 = *reinterpret_cast<const float4 *>(&sm.As[j][ty * 8 + 4]);
                const float4 b = *reinterpret_cast<const float4 *>(&sm.Bs[j][tx * 4]);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
            }
            __syncthreads();
        }

        // candidates that beat the user's current k-th best go to the user's queue
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int u = ty * 8 + a;
            const float th = sm.thr[u];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int64_t item = it0 + tx * 4 + c;
                if (item < it_end && acc[a][c] > th) {
                    const int slot = atomicAdd(&sm.q_cnt[u], 1);
                    sm.q_s[u][slot] = acc[a][c];
                    sm.q_i[u][slot] = (int)item;
                }
            }
        }
        __syncthreads();

        // merge: warp w owns users w*8 .. w*8+7; lane i holds the i-th best entry
        for (int uu = 0; uu < 8; ++uu) {
            const int u = warp * 8 + uu;
            const int cnt = sm.q_cnt[u];
            if (cnt == 0) continue;
            const int64_t q = q0 + u;
            float es = sm.tk_s[u][lane];
            int ei = sm.tk_i[u][lane];
            int64_t mb = 0, me = 0;
            if (mask_rowptr && q < nu) { mb = mask_rowptr[q]; me = mask_rowptr[q + 1]; }
            for (int c = 0; c < cnt; ++c) {
                const float cs = sm.q_s[u][c];
                const int ci = sm.q_i[u][c];
                // excluded (train) item?  uniform binary search in the user's sorted list
                int64_t lo = mb, hi = me;
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (__ldg(mask_col + mid) < ci) lo = mid + 1; else hi = mid;
                }
                if (lo < me && __ldg(mask_col + lo) == ci) continue;
                const bool b = (lane < k) && better(cs, ci, es, ei);
                const unsigned ball = __ballot_sync(0xffffffffu, b);
                if (ball == 0) continue;
                const int posn = __ffs(ball) - 1;            // first entry the candidate beats
                const float up_s = __shfl_up_sync(0xffffffffu, es, 1);
                const int up_i = __shfl_up_sync(0xffffffffu, ei, 1);
                if (lane == posn) { es = cs; ei = ci; }
                else if (lane > posn) { es = up_s; ei = up_i; }
            }
            sm.tk_s[u][lane] = es;
            sm.tk_i[u][lane] = ei;
            if (lane == k - 1) sm.thr[u] = es;
            if (lane == 0) sm.q_cnt[u] = 0;
        }
        __syncthreads();
    }

    for (int idx = tid; idx < TU * k; idx += kScoreThreads) {
        const int u = idx / k, r = idx % k;
        const int64_t q = q0 + u;
        if (q < nu) {
            out_ids[q * k + r] = sm.tk_i[u][r];
            out_scores[q * k + r] = sm.tk_s[u][r];
        }
    }
}

// merge the per-split partial top-k lists of one user (exact scores: the merge is exact)
__global__ void __launch_bounds__(256)
merge_topk_kernel(const int32_t *__restrict__ p_ids, const float *__restrict__ p_sc, int64_t nu, int k,
                  int n_splits, int32_t *__restrict__ out_ids, float *__restrict__ out_scores) {
    const int lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nu) return;
    const int total = n_splits * k;
    // every split list is sorted best-first: lane l walks lists l, l+32, ... with a cursor each
    constexpr int MAXL = 4;                       // n_splits <= 128
    int cur[MAXL];
#pragma unroll
    for (int j = 0; j < MAXL; ++j) cur[j] = 0;
    (void)total;
    for (int r = 0; r < k; ++r) {
        float bs = -FLT_MAX;
        int bi = -1, bj = -1;
#pragma unroll
        for (int j = 0; j < MAXL; ++j) {
            const int sp = lane + 32 * j;
            if (sp < n_splits && cur[j] < k) {
                const size_t o = ((size_t)sp * nu + q) * k + cur[j];
                const float s = p_sc[o];
                const int id = p_ids[o];
                if (id >= 0 && (bj < 0 || better(s, id, bs, bi))) { bs = s; bi = id; bj = j; }
            }
        }
        float ws = bs;
        int wi = bi, wl = lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, ws, off);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, off);
            const int ol = __shfl_xor_sync(0xffffffffu, wl, off);
            const bool take = oi >= 0 && (wi < 0 || better(os, oi, ws, wi));
            if (take) { ws = os; wi = oi; wl = ol; }
        }
        if (lane == 0) { out_ids[q * k + r] = wi; out_scores[q * k + r] = wi >= 0 ? ws : -FLT_MAX; }
        if (lane == wl && bj >= 0 && wi >= 0) {
#pragma unroll
            for (int j = 0; j < MAXL; ++j) if (j == bj) ++cur[j];
        }
    }
}

// hits and DCG of reference main.py:430-438: one thread per evaluated user
__global__ void eval_metrics_kernel(const int32_t *__restrict__ ids, const int64_t *__restrict__ targets,
                                    int64_t nu, int k, double *__restrict__ sums) {
    double hit = 0.0, dcg = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nu; q += stride) {
        const int64_t t = targets[q];
        for (int r = 0; r < k; ++r)
            if ((int64_t)ids[q * k + r] == t) {
                hit += 1.0;
                dcg += 1.0 / log2((double)r + 2.0);
                break;
            }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        hit += __shfl_xor_sync(0xffffffffu, hit, off);
        dcg += __shfl_xor_sync(0xffffffffu, dcg, off);
    }
    if ((threadIdx.x & 31) == 0 && (hit != 0.0 || dcg != 0.0)) {
        atomicAdd(sums, hit);
        atomicAdd(sums + 1, dcg);
    }
}

}  // namespace lgcn

namespace {
// item splits so that a small number of user tiles still fills the chip
int choose_splits(int64_t nu, int64_t n_items) {
    const int64_t user_tiles = (nu + lgcn::TU - 1) / lgcn::TU;
    if (user_tiles >= 2 * lgcn::kNumSMs || n_items < 64 * lgcn::TI) return 1;
    int64_t s = (2 * lgcn::kNumSMs + user_tiles - 1) / user_tiles;
    const int64_t max_by_items = n_items / (16 * lgcn::TI);
    if (s > max_by_items) s = max_by_items;
    if (s > 128) s = 128;
    return s < 1 ? 1 : (int)s;
}
}  // namespace

extern "C" size_t lgcn_score_topk_workspace(int64_t nu, int64_t n_items, int32_t d, int32_t k) {
    (void)d;
    if (nu <= 0 || n_items <= 0 || k <= 0) return 0;
    const int s = choose_splits(nu, n_items);
    return s <= 1 ? 0 : (size_t)s * nu * k * 8;
}

extern "C" int lgcn_score_topk(const float *Fu, const float *Fi, const int64_t *users, int64_t nu,
                               int64_t n_items, int32_t d, const int64_t *mask_rowptr,
                               const int32_t *mask_col, int32_t k, int32_t *out_ids,
                               float *out_scores, void *workspace, size_t workspace_bytes,
                               lgcn_stream_t stream) {
    using namespace lgcn;
    if (d <= 0 || d % 4 != 0) return LGCN_E_BAD_DIM;
    if (nu < 0 || n_items <= 0 || k <= 0 || k > KMAX || !Fu || !Fi || !out_ids || !out_scores)
        return LGCN_E_BAD_ARG;
    if (nu == 0) return 0;
    if (!users) return LGCN_E_BAD_ARG;
    if (mask_rowptr && !mask_col) return LGCN_E_BAD_ARG;
    if (n_items > 0x7fffffffLL) return LGCN_E_TOO_LARGE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    LGCN_OPT_IN_SMEM(score_topk_kernel, sizeof(ScoreSmem));
    const int64_t blocks = (nu + TU - 1) / TU;
    if (blocks > 0x7fffffffLL) return LGCN_E_TOO_LARGE;
    int splits = choose_splits(nu, n_items);
    if (splits > 1 && (!workspace || workspace_bytes < (size_t)splits * nu * k * 8)) splits = 1;
    if (splits == 1) {
        score_topk_kernel<<<(unsigned)blocks, kScoreThreads, sizeof(ScoreSmem), st>>>(
            Fu, Fi, users, nu, n_items, d, mask_rowptr, mask_col, k, out_ids, out_scores, n_items);
        LGCN_LAUNCH_CHECK();
        return 0;
    }
    int64_t per = (n_items + splits - 1) / splits;
    per = (per + TI - 1) / TI * TI;
    splits = (int)((n_items + per - 1) / per);
    int32_t *p_ids = reinterpret_cast<int32_t *>(workspace);
    float *p_sc = reinterpret_cast<float *>(p_ids + (size_t)splits * nu * k);
    dim3 grid((unsigned)blocks, (unsigned)splits);
    score_topk_kernel<<<grid, kScoreThreads, sizeof(ScoreSmem), st>>>(
        Fu, Fi, users, nu, n_items, d, mask_rowptr, mask_col, k, p_ids, p_sc, per);
    LGCN_LAUNCH_CHECK();
    merge_topk_kernel<<<(unsigned)((nu * 32 + 255) / 256), 256, 0, st>>>(p_ids, p_sc, nu, k, splits, out_ids,
                                                                           out_scores);
    LGCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int lgcn_eval_metrics(const int32_t *topk_ids, const int64_t *targets, int64_t nu,
                                 int32_t k, double *sums, lgcn_stream_t stream) {
    if (nu < 0 || k <= 0 || !topk_ids || !targets || !sums) return LGCN_E_BAD_ARG;
    if (nu == 0) return 0;
    int64_t blocks = (nu + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    lgcn::eval_metrics_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        topk_ids, targets, nu, k, sums);
    LGCN_LAUNCH_CHECK();
    return 0;
}
