// lgcn_score_tc.cu -- full-rank rating on the 5th-gen tensor cores (sm_100a): tcgen05.mma with
// TMEM accumulators, item tiles streamed by bulk async copies (TMA unit), train-item masking and
// a per-user candidate heap fused into the epilogue, then an exact fp32 re-score.
//
// Replaces torch.matmul(user_batch, item_table.T) + mask loop + torch.topk of reference
// main.py:420-426 for large catalogues.  Bit-exact ids need fp32 scores, tensor cores take
// bf16: so this is FILTER-AND-REFINE (hard part 2):
//
//  1. lgcn_score_tc_prepare : item table fp32 -> bf16 in the UMMA canonical K-major layout, one
//                             contiguous 128-item tile per bulk copy; max item norm.
//  2. score_filter_kernel   : CTA = 128 users.  S[128 x 128] = U_bf16 . V_bf16^T per item tile
//                             (d/16 tcgen05.mma of 128x128x16, fp32 accumulate in TMEM, two
//                             accumulator buffers).  Epilogue threads own one user (= one TMEM
//                             lane) each: tcgen05.ld 32 scores at a time, skip the user's train
//                             items with a cursor into the sorted mask list, and push scores
//                             above the user's running threshold into a C-entry min-heap in
//                             shared memory.  The [users x items] scores never leave the SM.
//                             The heap key is not the bf16 score s but an UPPER BOUND of the exact
//                             score: key = s + c_u * |v|, c_u = 1.05 * 2^-8 * |u| (bf16 rounding
//                             of both operands: |s - u.v| <= (2^-8 + 2^-16) |u||v|), with the
//                             item's own norm -- a window of 32 items is first tested against
//                             its largest norm (one FMA per window), so the bound costs nothing
//                             in the common case and is not inflated by the few high-norm items.
//  3. score_refine_kernel   : exact fp32 score (sequential FMA over the features, the order of
//                             the CPU oracle) of the C candidates, ordered top-k, and a per-user
//                             certificate: every dropped item has exact <= key <= threshold, so
//                             exact_kth > threshold proves the ids; users that fail are flagged
//                             and re-run by the exact SIMT kernel (lgcn_score.cu).
#include <cuda_bf16.h>
#include <float.h>

#include "lgcn_common.cuh"

namespace lgcn {
namespace tc {

constexpr int MT = 128;        // users per CTA (UMMA M)
constexpr int NT = 128;        // items per tile (UMMA N)
#ifndef LGCN_TC_NSTAGE
#define LGCN_TC_NSTAGE 3
#endif
#ifndef LGCN_TC_CAND
#define LGCN_TC_CAND 96
#endif
constexpr int NSTAGE = LGCN_TC_NSTAGE;   // item-tile ring
constexpr int CAND = LGCN_TC_CAND;       // candidates kept per user (all epilogue groups together)
constexpr int NB = 4;          // TMEM accumulator buffers (tile t -> accumulator t % NB) == candidate heaps per user
constexpr int GCAND = CAND / NB;             // heap entries per (user, epilogue group, column half)
constexpr int kThreads = 64 + NB * 128;      // warp 0: copy producer, warp 1: MMA issuer, then two
                                             // epilogue groups of 8 warps (2 column halves x 4 TMEM lane quadrants)
static_assert(NB == 4, "two epilogue groups x two accumulators each");
constexpr int TMEM_COLS = NB * 128;          // NB 128-column fp32 accumulators (all 512 columns)

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, kind::f16 (bf16 in, fp32 accumulate)
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (asynchronous: the values
// are valid after tc_ld_wait)
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, no swizzle (canonical "interleaved" layout):
// core matrix = 8 rows x 16 bytes stored contiguously (128 B); SBO = byte distance between 8-row
// groups, LBO = byte distance between the two 16-byte K chunks of one K=16 instruction.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    return d;                               // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M=128, N=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) |
                            ((uint32_t)(MT >> 4) << 24);

// element offset of (row r, feature k) inside a canonical 128-row tile of D features
__host__ __device__ inline int canon_off(int r, int k) {
    return (k >> 3) * (16 * 64) + (r >> 3) * 64 + (r & 7) * 8 + (k & 7);
}

template <int D>
struct FilterSmem {
    __nv_bfloat16 A[MT * D];
    __nv_bfloat16 B[NSTAGE][NT * D];
    float heap_s[NB][GCAND][MT];
    int heap_i[NB][GCAND][MT];
    unsigned long long full[NSTAGE], empty[NSTAGE], tfull[NB], tempty[NB];
    uint32_t tmem_base;
};

// ---- item table -> bf16 canonical tiles, item norms -----------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
prepare_items_kernel(const float *__restrict__ Fi, int64_t n_items, int64_t n_pad,
                     __nv_bfloat16 *__restrict__ Bt, float *__restrict__ vnorm) {
    // one thread per (item, 8-feature chunk)
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int KCH = D / 8;
    if (idx >= n_pad * KCH) return;
    const int64_t item = idx / KCH;
    const int kc = (int)(idx % KCH);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (item < n_items) {
        a = ld_stream_f4(Fi + item * D + kc * 8);
        b = ld_stream_f4(Fi + item * D + kc * 8 + 4);
    }
    __nv_bfloat162 h[4];
    h[0] = __floats2bfloat162_rn(a.x, a.y);
    h[1] = __floats2bfloat162_rn(a.z, a.w);
    h[2] = __floats2bfloat162_rn(b.x, b.y);
    h[3] = __floats2bfloat162_rn(b.z, b.w);
    const int64_t tile = item / NT;
    const int r = (int)(item % NT);
    __nv_bfloat16 *dst = Bt + tile * (int64_t)(NT * D) + canon_off(r, kc * 8);
    *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(h);
    // squared norm of the chunk, reduced over the KCH threads of the item (consecutive lanes)
    float s = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
#pragma unroll
    for (int off = KCH / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    // rounded up by one ulp-ish factor so that it is an upper bound of the true norm
    if (kc == 0) vnorm[item] = sqrtf(s) * 1.000001f;
}

// largest item norm of every 32-item window (the granularity of the epilogue's tcgen05.ld)
__global__ void __launch_bounds__(256)
window_norm_kernel(const float *__restrict__ vnorm, int64_t n_windows, float *__restrict__ wnorm) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_windows) return;
    const float4 *p = reinterpret_cast<const float4 *>(vnorm + w * 32);
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 v = __ldg(p + i);
        m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
    wnorm[w] = m;
}

// ---- heap helpers: min-heap on (score asc, id desc) so the root is the WORST kept candidate ----
__device__ __forceinline__ bool worse(float s, int id, float t, int tid) {
    return s < t || (s == t && id > tid);
}

// Replace the root (worst kept candidate) of user u's min-heap and sift down.  NOT inlined: it is
// called from 32 sites per 32-column window and runs ~1e-4 of the time; inlining it made the
// epilogue loop ~200 KB of code and instruction-cache bound (measured: 9200 cycles per tile).
__device__ __noinline__ float heap_push(float *heap_s, int *heap_i, int u, float s, int id) {
    int pos = 0;
    while (true) {
        const int l = 2 * pos + 1;
        if (l >= GCAND) break;
        int c = l;
        float cs = heap_s[l * MT + u];
        int ci = heap_i[l * MT + u];
        if (l + 1 < GCAND) {
            const float rs = heap_s[(l + 1) * MT + u];
            const int ri = heap_i[(l + 1) * MT + u];
            if (worse(rs, ri, cs, ci)) { c = l + 1; cs = rs; ci = ri; }
        }
        if (!worse(cs, ci, s, id)) break;
        heap_s[pos * MT + u] = cs;
        heap_i[pos * MT + u] = ci;
        pos = c;
    }
    heap_s[pos * MT + u] = s;
    heap_i[pos * MT + u] = id;
    return heap_s[u];                     // new threshold = score of the root
}

// ---- tensor-core filter ------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kThreads, 1)
score_filter_kernel(const float *__restrict__ Fu, const int64_t *__restrict__ users, int64_t nu,
                    const __nv_bfloat16 *__restrict__ Bt, const float *__restrict__ vnorm,
                    const float *__restrict__ wnorm, int64_t n_items, int n_tiles_all, int tiles_per_split,
                    const int64_t *__restrict__ mask_rowptr, const int32_t *__restrict__ mask_col,
                    float *__restrict__ cand_s, int32_t *__restrict__ cand_i, float *__restrict__ tau_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FilterSmem<D> &sm = *reinterpret_cast<FilterSmem<D> *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t q0 = (int64_t)blockIdx.x * MT;
    // blockIdx.y = item split (few users: the catalogue is cut so that the chip is filled): this
    // CTA sweeps item tiles [t0, t0 + n_tiles) and keeps its own candidate set per user
    const int t0 = (int)blockIdx.y * tiles_per_split;
    const int n_tiles = min(tiles_per_split, n_tiles_all - t0);
    cand_s += (size_t)blockIdx.y * nu * CAND;
    cand_i += (size_t)blockIdx.y * nu * CAND;
    tau_out += (size_t)blockIdx.y * nu * NB;
    constexpr uint32_t TILE_BYTES = NT * D * 2;
    constexpr uint32_t LBO = 16 * 128;     // bytes between K chunks (16 row groups of 128 B)
    constexpr uint32_t SBO = 128;          // bytes between 8-row groups

    // ---- one-time setup ------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(smem_u32(&sm.full[s]), 1); mbar_init(smem_u32(&sm.empty[s]), 1); }
        for (int b = 0; b < NB; ++b) { mbar_init(smem_u32(&sm.tfull[b]), 1); mbar_init(smem_u32(&sm.tempty[b]), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&sm.tmem_base)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // user tile: fp32 rows -> bf16, canonical K-major layout (generic-proxy stores)
    for (int ch = tid; ch < MT * (D / 8); ch += kThreads) {
        const int m = ch / (D / 8), kc = ch % (D / 8);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (q0 + m < nu) {
            const float *src = Fu + (size_t)users[q0 + m] * D + kc * 8;
            a = ld_nc_f4(src);
            b = ld_nc_f4(src + 4);
        }
        __nv_bfloat162 h[4];
        h[0] = __floats2bfloat162_rn(a.x, a.y);
        h[1] = __floats2bfloat162_rn(a.z, a.w);
        h[2] = __floats2bfloat162_rn(b.x, b.y);
        h[3] = __floats2bfloat162_rn(b.z, b.w);
        *reinterpret_cast<uint4 *>(&sm.A[canon_off(m, kc * 8)]) = *reinterpret_cast<const uint4 *>(h);
    }
    for (int i = tid; i < CAND * MT; i += kThreads) {
        (&sm.heap_s[0][0][0])[i] = -FLT_MAX;
        (&sm.heap_i[0][0][0])[i] = 0x7fffffff;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared -> visible to UMMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp == 0) {
        // ===== producer: one bulk async copy (TMA unit) per 128-item tile =====
        if (lane == 0) {
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % NSTAGE;
                mbar_wait(smem_u32(&sm.empty[s]), ((t / NSTAGE) & 1) ^ 1);
                mbar_expect_tx(smem_u32(&sm.full[s]), TILE_BYTES);
                bulk_g2s(smem_u32(&sm.B[s][0]), Bt + (size_t)(t0 + t) * (NT * D), TILE_BYTES, smem_u32(&sm.full[s]));
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: a single thread drives the tensor core =====
        if (lane == 0) {
            const uint32_t a_base = smem_u32(&sm.A[0]);
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % NSTAGE, b = t % NB;
                mbar_wait(smem_u32(&sm.tempty[b]), ((t / NB) & 1) ^ 1);    // epilogue drained this accumulator
                mbar_wait(smem_u32(&sm.full[s]), (t / NSTAGE) & 1);        // tile landed in smem
                tc_fence_after();
                const uint32_t b_base = smem_u32(&sm.B[s][0]);
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint64_t ad = make_smem_desc(a_base + kk * 2 * LBO, LBO, SBO);
                    const uint64_t bd = make_smem_desc(b_base + kk * 2 * LBO, LBO, SBO);
                    tc_mma(tmem_base + b * NT, ad, bd, kIdesc, kk > 0 ? 1u : 0u);
                }
                tc_commit(smem_u32(&sm.empty[s]));    // smem slot reusable once these MMAs retire
                tc_commit(smem_u32(&sm.tfull[b]));    // accumulator ready for the epilogue
            }
        }
    } else {
        // ===== epilogue: two groups of 8 warps; group g owns tiles t = g (mod 2) and therefore TWO
        // accumulators (t mod 4) that it fills and drains alternately: while it works on tile t the
        // MMA already runs into its other accumulator (tile t + 2), so neither side waits out the
        // commit -> wake-up -> arrive round trip (ncu r02 on the one-accumulator-per-group design:
        // the MMA thread never waited for a tile to land, it waited for `tempty`, and the epilogue
        // warps spun on `tfull` half of the time -- a latency-bound handshake, tensor pipe 32 %).
        // Inside a group the tile's 128 columns are split between two warps per TMEM lane quadrant
        // (thread <-> user <-> lane; half h takes columns [64h, 64h + 64)), each with its own heap.
        const int ew = warp - 2;
        const int g = ew >> 3;                        // epilogue group: tiles t = g (mod 2)
        const int half = (ew >> 2) & 1;               // column half of the tile
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
        const int hsel = g * 2 + half;                // heap / threshold slot of this (group, half)
        const int u = quad * 32 + lane;
        const int64_t q = q0 + u;
        float *hs = &sm.heap_s[hsel][0][0];
        int *hi = &sm.heap_i[hsel][0][0];
        float tau = -FLT_MAX;
        // c_u = 1.05 * 2^-8 * |u|: key = s + c_u |v| bounds the exact score from above
        float cu = 0.f;
        if (q < nu) {
            const float4 *fu4 = reinterpret_cast<const float4 *>(Fu + (size_t)users[q] * D);
            float n2 = 0.f;
#pragma unroll 4
            for (int j = 0; j < D / 4; ++j) {
                const float4 a = __ldg(fu4 + j);
                n2 += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
            }
            cu = 1.05f * 0.00390625f * sqrtf(n2) * 1.000001f;
        }
        // window norms: two 32-item windows per (tile, half)
        const float2 *wn2 = reinterpret_cast<const float2 *>(wnorm);
        float2 wn_next = g < n_tiles ? __ldg(wn2 + (size_t)(t0 + g) * 2 + half) : make_float2(0.f, 0.f);
        int64_t mb = 0, me = 0;
        if (mask_rowptr && q < nu) { mb = mask_rowptr[q]; me = mask_rowptr[q + 1]; }
        int next_masked = (mb < me) ? __ldg(mask_col + mb) : 0x7fffffff;
        for (int t = g; t < n_tiles; t += 2) {
            const int b = t % NB;                          // accumulator of this tile
            const float2 wn_t = wn_next;                   // window norms of this half, fetched a tile ahead
            if (t + 2 < n_tiles) wn_next = __ldg(wn2 + (size_t)(t0 + t + 2) * 2 + half);
            mbar_wait(smem_u32(&sm.tfull[b]), ((t / NB) & 1));
            tc_fence_after();
            const int half_item0 = (t0 + t) * NT + half * 64;
            // this warp sees one half of every second tile: advance the mask cursor to its start
            while (next_masked < half_item0) {            // register compare; loads only on advance
                ++mb;
                next_masked = (mb < me) ? __ldg(mask_col + mb) : 0x7fffffff;
            }
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t raw[32];
                tc_ld32_issue(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * NT + half * 64 + c * 32), raw);
                tc_ld_wait();
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                const int item0 = half_item0 + c * 32;
                // the user's train items inside this 32-column window (sorted list, cursor)
                while (next_masked < item0 + 32) {
                    const int j = next_masked - item0;
#pragma unroll
                    for (int i = 0; i < 32; ++i) if (i == j) v[i] = -FLT_MAX;
                    ++mb;
                    next_masked = (mb < me) ? __ldg(mask_col + mb) : 0x7fffffff;
                }
                if ((int64_t)item0 + 32 > n_items) {          // zero padding of the last tile
#pragma unroll
                    for (int i = 0; i < 32; ++i) if ((int64_t)item0 + i >= n_items) v[i] = -FLT_MAX;
                }
                // window maximum as a balanced tree (a 31-deep FMNMX chain is latency bound)
                float m16[16], m8[8], m4[4];
#pragma unroll
                for (int i = 0; i < 16; ++i) m16[i] = fmaxf(v[2 * i], v[2 * i + 1]);
#pragma unroll
                for (int i = 0; i < 8; ++i) m8[i] = fmaxf(m16[2 * i], m16[2 * i + 1]);
#pragma unroll
                for (int i = 0; i < 4; ++i) m4[i] = fmaxf(m8[2 * i], m8[2 * i + 1]);
                const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                const float wn = c == 0 ? wn_t.x : wn_t.y;
                // Rare for one thread once its heap has warmed up -- but a tile is released only when
                // all of its 8 warps are through, and with ~10 % of the warp-windows in here most
                // tiles wait for one (ncu r02): keep this path free of loads.  A score can only enter
                // the heap if it beats the threshold even with the window's LARGEST norm, so the
                // item's own norm is fetched for those few scores alone (its key <= that bound).
                if (fmaf(cu, wn, mx) > tau) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (fmaf(cu, wn, v[i]) > tau) {
                            const float key = fmaf(cu, __ldg(vnorm + item0 + i), v[i]);
                            if (key > tau) tau = heap_push(hs, hi, u, key, item0 + i);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sm.tempty[b]));
        }
        if (q < nu) {
            for (int k = 0; k < GCAND; ++k) {
                cand_s[q * CAND + hsel * GCAND + k] = hs[k * MT + u];
                cand_i[q * CAND + hsel * GCAND + k] = hi[k * MT + u];
            }
            tau_out[q * NB + hsel] = tau;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ---- exact re-score of the candidates, ordered top-k, certificate ------------------------------
// one warp per user; lane owns candidates lane and lane+32
__global__ void __launch_bounds__(256)
score_refine_kernel(const float *__restrict__ Fu, const float *__restrict__ Fi,
                    const int64_t *__restrict__ users, int64_t nu, int d,
                    const int32_t *__restrict__ cand_i, const float *__restrict__ tau,
                    int k, int32_t *__restrict__ out_ids,
                    float *__restrict__ out_scores, int32_t *__restrict__ fail) {
    // blockIdx.y = item split: candidates, thresholds and the ordered top-k list of (split, user)
    // live at index split*nu + user; with several splits the lists are merged (and certified) by
    // merge_cert_kernel and `fail` is NULL here
    const int lane = threadIdx.x & 31;
    const int64_t qu = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (qu >= nu) return;
    const float *fu = Fu + (size_t)users[qu] * d;
    const int64_t q = (int64_t)blockIdx.y * nu + qu;
    constexpr int CPL = CAND / 32;                 // candidates per lane
    float sc[CPL];
    int id[CPL];
#pragma unroll
    for (int h = 0; h < CPL; ++h) {
        id[h] = cand_i[q * CAND + lane + 32 * h];
        sc[h] = -FLT_MAX;
        if (id[h] != 0x7fffffff) {
            const float *fi = Fi + (size_t)id[h] * d;
            float acc = 0.f;
            for (int j = 0; j < d; j += 4) {          // sequential fp32 FMA over the features
                const float4 a = *reinterpret_cast<const float4 *>(fu + j);
                const float4 b = ld_nc_f4(fi + j);
                acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
                acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
            }
            sc[h] = acc;
        } else {
            id[h] = -1;
        }
    }
    // k rounds of warp arg-best on (score desc, id asc)
    float kth = -FLT_MAX;
    for (int r = 0; r < k; ++r) {
        int hb = 0;
#pragma unroll
        for (int h = 1; h < CPL; ++h)
            if (sc[h] > sc[hb] || (sc[h] == sc[hb] && (unsigned)id[h] < (unsigned)id[hb])) hb = h;
        float bs = sc[0];
        int bi = id[0];
#pragma unroll
        for (int h = 1; h < CPL; ++h) if (h == hb) { bs = sc[h]; bi = id[h]; }
        int bl = lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            const int ol = __shfl_xor_sync(0xffffffffu, bl, off);
            if (os > bs || (os == bs && (unsigned)oi < (unsigned)bi)) { bs = os; bi = oi; bl = ol; }
        }
        if (lane == 0) { out_ids[q * k + r] = bi; out_scores[q * k + r] = bs; }
        if (lane == bl) {                                   // id -1 sorts last among -FLT_MAX ties
#pragma unroll
            for (int h = 0; h < CPL; ++h) if (h == hb) { sc[h] = -FLT_MAX; id[h] = -1; }
        }
        kth = bs;
    }
    if (lane == 0 && fail) {
        // every item the filter dropped has exact score <= key <= its group's threshold (the keys
        // are upper bounds of the exact scores): the k-th exact score must clear the largest one.
        float tq = tau[q * NB];
#pragma unroll
        for (int g = 1; g < NB; ++g) tq = fmaxf(tq, tau[q * NB + g]);
        fail[q] = (kth > tq) ? 0 : 1;
    }
}

// Item-split runs: merge the per-split ordered top-k lists of one user (exact scores, so the merge
// is exact; (score desc, id asc) order) and certify against the largest threshold of ANY split /
// epilogue group: every dropped item has exact <= key <= its own (split, group) threshold.
__global__ void __launch_bounds__(256)
merge_cert_kernel(const int32_t *__restrict__ p_ids, const float *__restrict__ p_sc,
                  const float *__restrict__ tau, int64_t nu, int k, int n_splits,
                  int32_t *__restrict__ out_ids, float *__restrict__ out_scores, int32_t *__restrict__ fail) {
    const int lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nu) return;
    constexpr int MAXL = 2;                       // n_splits <= 64: lane l walks lists l and l+32
    int cur[MAXL] = {0, 0};
    float kth = -FLT_MAX;
    for (int r = 0; r < k; ++r) {
        float bs = -FLT_MAX;
        int bi = -1, bj = -1;
#pragma unroll
        for (int j = 0; j < MAXL; ++j) {
            const int sp = lane + 32 * j;
            if (sp < n_splits && cur[j] < k) {
                const size_t o = ((size_t)sp * nu + q) * k + cur[j];
                const float sc = p_sc[o];
                const int id = p_ids[o];
                if (id >= 0 && (bj < 0 || sc > bs || (sc == bs && id < bi))) { bs = sc; bi = id; bj = j; }
            }
        }
        float ws = bs;
        int wi = bi, wl = lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, ws, off);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, off);
            const int ol = __shfl_xor_sync(0xffffffffu, wl, off);
            if (oi >= 0 && (wi < 0 || os > ws || (os == ws && oi < wi))) { ws = os; wi = oi; wl = ol; }
        }
        if (lane == 0) { out_ids[q * k + r] = wi; out_scores[q * k + r] = wi >= 0 ? ws : -FLT_MAX; }
        if (lane == wl && bj >= 0 && wi >= 0) {
#pragma unroll
            for (int j = 0; j < MAXL; ++j) if (j == bj) ++cur[j];
        }
        kth = wi >= 0 ? ws : -FLT_MAX;
    }
    float tq = -FLT_MAX;
    for (int i = lane; i < n_splits * NB; i += 32) {
        const int sp = i / NB, g = i % NB;
        tq = fmaxf(tq, tau[((size_t)sp * nu + q) * NB + g]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tq = fmaxf(tq, __shfl_xor_sync(0xffffffffu, tq, off));
    if (lane == 0) fail[q] = (kth > tq) ? 0 : 1;
}

}  // namespace tc
}  // namespace lgcn

// ---- C ABI ---------------------------------------------------------------------------------
namespace {
// item splits so that a handful of user tiles (the reference rates 1024 users per batch,
// main.py:415) still fills the chip: user_tiles * splits <= one wave of CTAs
int tc_splits(int64_t nu, int64_t n_items) {
    using namespace lgcn::tc;
    const int64_t ut = (nu + MT - 1) / MT;
    const int64_t n_tiles = (n_items + NT - 1) / NT;
    if (ut <= 0 || ut >= lgcn::kNumSMs) return 1;
    int64_t s = lgcn::kNumSMs / ut;
    if (s > n_tiles / 32) s = n_tiles / 32;           // >= 32 tiles per split
    if (s > 64) s = 64;
    return s < 1 ? 1 : (int)s;
}

struct TcLayout {
    size_t bt, vnorm, wnorm, cand_s, cand_i, tau, p_ids, p_sc, total;
    int splits;
};
TcLayout tc_layout(int64_t nu, int64_t n_items, int32_t d) {
    using namespace lgcn::tc;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    TcLayout L;
    const int64_t n_pad = (n_items + NT - 1) / NT * NT;
    L.splits = tc_splits(nu, n_items);
    const size_t rows = (size_t)nu * L.splits;          // (split, user) pairs
    L.bt = 0;                                           // prepared item tiles (independent of nu)
    L.vnorm = up(L.bt + (size_t)n_pad * d * 2);         // item norms (independent of nu)
    L.wnorm = up(L.vnorm + (size_t)n_pad * 4);          // largest norm per 32-item window
    L.cand_s = up(L.wnorm + (size_t)(n_pad / 32) * 4);
    L.cand_i = up(L.cand_s + rows * CAND * 4);
    L.tau = up(L.cand_i + rows * CAND * 4);
    L.p_ids = up(L.tau + rows * 4 * lgcn::tc::NB);      // per-split ordered top-k (splits > 1)
    L.p_sc = up(L.p_ids + (L.splits > 1 ? rows * 32 * 4 : 0));
    L.total = up(L.p_sc + (L.splits > 1 ? rows * 32 * 4 : 0));
    return L;
}
}  // namespace

// The workspace serves every call with AT MOST nu users (a sweep reuses one workspace for all of
// its batches, the last one shorter): fewer users mean more item splits, so take the largest.
extern "C" size_t lgcn_score_tc_workspace(int64_t nu, int64_t n_items, int32_t d) {
    if (nu < 0 || n_items <= 0 || (d != 64 && d != 128)) return 0;
    size_t best = tc_layout(nu, n_items, d).total;
    for (int64_t ut = 1; ut < lgcn::kNumSMs; ++ut) {
        const int64_t n = ut * lgcn::tc::MT;
        if (n >= nu) break;
        const size_t t = tc_layout(n, n_items, d).total;
        if (t > best) best = t;
    }
    return best;
}

extern "C" int lgcn_score_tc_prepare(const float *Fi, int64_t n_items, int32_t d, void *workspace,
                                     size_t workspace_bytes, lgcn_stream_t stream) {
    using namespace lgcn::tc;
    if (d != 64 && d != 128) return LGCN_E_BAD_DIM;
    if (!Fi || n_items <= 0 || !workspace) return LGCN_E_BAD_ARG;
    const TcLayout L = tc_layout(0, n_items, d);
    if (workspace_bytes < L.cand_s) return LGCN_E_BAD_ARG;
    if (n_items > 0x7fffff00LL) return LGCN_E_TOO_LARGE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char *ws = reinterpret_cast<unsigned char *>(workspace);
    const int64_t n_pad = (n_items + NT - 1) / NT * NT;
    float *vnorm = reinterpret_cast<float *>(ws + L.vnorm);
    float *wnorm = reinterpret_cast<float *>(ws + L.wnorm);
    const int64_t threads = n_pad * (d / 8);
    const unsigned grid = (unsigned)((threads + 255) / 256);
    if (d == 64) prepare_items_kernel<64><<<grid, 256, 0, st>>>(Fi, n_items, n_pad, reinterpret_cast<__nv_bfloat16 *>(ws + L.bt), vnorm);
    else prepare_items_kernel<128><<<grid, 256, 0, st>>>(Fi, n_items, n_pad, reinterpret_cast<__nv_bfloat16 *>(ws + L.bt), vnorm);
    LGCN_LAUNCH_CHECK();
    const int64_t n_windows = n_pad / 32;
    window_norm_kernel<<<(unsigned)((n_windows + 255) / 256), 256, 0, st>>>(vnorm, n_windows, wnorm);
    LGCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int lgcn_score_tc_topk(const float *Fu, const float *Fi, const int64_t *users, int64_t nu,
                                  int64_t n_items, int32_t d, const int64_t *mask_rowptr,
                                  const int32_t *mask_col, int32_t k, int32_t *out_ids,
                                  float *out_scores, int32_t *fail, void *workspace,
                                  size_t workspace_bytes, lgcn_stream_t stream) {
    using namespace lgcn::tc;
    if (d != 64 && d != 128) return LGCN_E_BAD_DIM;
    if (nu < 0 || n_items <= 0 || k <= 0 || k > 32 || !Fu || !Fi || !out_ids || !out_scores || !fail || !workspace)
        return LGCN_E_BAD_ARG;
    if (nu == 0) return 0;
    if (!users || (mask_rowptr && !mask_col)) return LGCN_E_BAD_ARG;
    const TcLayout L = tc_layout(nu, n_items, d);
    if (workspace_bytes < L.total) return LGCN_E_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char *ws = reinterpret_cast<unsigned char *>(workspace);
    const __nv_bfloat16 *Bt = reinterpret_cast<const __nv_bfloat16 *>(ws + L.bt);
    float *cand_s = reinterpret_cast<float *>(ws + L.cand_s);
    int32_t *cand_i = reinterpret_cast<int32_t *>(ws + L.cand_i);
    float *tau = reinterpret_cast<float *>(ws + L.tau);
    const float *vnorm = reinterpret_cast<const float *>(ws + L.vnorm);
    const float *wnorm = reinterpret_cast<const float *>(ws + L.wnorm);
    const int n_tiles = (int)((n_items + NT - 1) / NT);
    const int splits = L.splits;
    const int tiles_per_split = (n_tiles + splits - 1) / splits;
    const int splits_used = (n_tiles + tiles_per_split - 1) / tiles_per_split;
    const dim3 grid((unsigned)((nu + MT - 1) / MT), (unsigned)splits_used);
    if (d == 64) {
        LGCN_OPT_IN_SMEM((score_filter_kernel<64>), sizeof(FilterSmem<64>));
        score_filter_kernel<64><<<grid, kThreads, sizeof(FilterSmem<64>), st>>>(Fu, users, nu, Bt, vnorm, wnorm, n_items, n_tiles, tiles_per_split, mask_rowptr, mask_col, cand_s, cand_i, tau);
    } else {
        LGCN_OPT_IN_SMEM((score_filter_kernel<128>), sizeof(FilterSmem<128>));
        score_filter_kernel<128><<<grid, kThreads, sizeof(FilterSmem<128>), st>>>(Fu, users, nu, Bt, vnorm, wnorm, n_items, n_tiles, tiles_per_split, mask_rowptr, mask_col, cand_s, cand_i, tau);
    }
    LGCN_LAUNCH_CHECK();
    const dim3 rgrid((unsigned)((nu * 32 + 255) / 256), (unsigned)splits_used);
    if (splits_used == 1) {
        score_refine_kernel<<<rgrid, 256, 0, st>>>(Fu, Fi, users, nu, d, cand_i, tau, k, out_ids, out_scores, fail);
        LGCN_LAUNCH_CHECK();
        return 0;
    }
    int32_t *p_ids = reinterpret_cast<int32_t *>(ws + L.p_ids);
    float *p_sc = reinterpret_cast<float *>(ws + L.p_sc);
    score_refine_kernel<<<rgrid, 256, 0, st>>>(Fu, Fi, users, nu, d, cand_i, tau, k, p_ids, p_sc, nullptr);
    LGCN_LAUNCH_CHECK();
    merge_cert_kernel<<<(unsigned)((nu * 32 + 255) / 256), 256, 0, st>>>(p_ids, p_sc, tau, nu, k, splits_used,
                                                                           out_ids, out_scores, fail);
    LGCN_LAUNCH_CHECK();
    return 0;
}

// Host-only query: kernels one lgcn_score_tc_topk call launches (2 = filter + refine, 3 with the
// merge of an item-split run) -- for launch accounting.
extern "C" int lgcn_score_tc_launches(int64_t nu, int64_t n_items) {
    if (nu <= 0 || n_items <= 0) return 0;
    return tc_splits(nu, n_items) > 1 ? 3 : 2;
}
