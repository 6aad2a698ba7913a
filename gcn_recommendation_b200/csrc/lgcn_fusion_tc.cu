// lgcn_fusion_tc.cu -- LightGCN_Fusion item block on the 5th-gen tensor cores (sm_100a).
//
// Replaces reference models/lightgcn_fusion.py:45-49 (forward):
//     H = leaky_relu( cat([E_id, C], 1) @ W.T + b ),   W: [d, d+c],  c = 768
// fp32-faithful (<= 1e-5) on tensor cores by the 3xTF32 split: x = hi + lo with hi = x truncated
// to 10 mantissa bits (exact), lo = x - hi, and  x.w ~= hi.whi + lo.whi + hi.wlo  (the dropped
// lo.wlo term and the truncation of lo are ~2^-21 relative).  The concatenation is never
// materialised and the 13.5 GB content matrix is streamed exactly once.
//
// CTA (persistent over 128-item tiles): 16 loader warps read fp32 rows (E_id for k < d, content
// for k >= d; W rows likewise), split them and write hi / lo straight into the UMMA K-major
// SWIZZLE_128B layout in shared memory (generic-proxy stores + fence.proxy.async); one
// thread issues 12 tcgen05.mma.kind::tf32 (128 x d x 8) per 32-wide K chunk into one of two TMEM
// accumulators; 4 epilogue warps (thread <-> item <-> TMEM lane) add the bias, apply the leaky
// relu and store the row.  3-stage ring, mbarrier full/empty, tcgen05.commit.
#include <float.h>

#include "lgcn_common.cuh"

namespace lgcn {
namespace ftc {

constexpr int MT = 128;            // items per tile (UMMA M)
constexpr int KC = 32;             // K chunk (floats) = one 128-byte swizzle-atom row = 4 K=8 instructions
constexpr int NSTAGE = 3;
#ifndef LGCN_FW_LOADER_WARPS
#define LGCN_FW_LOADER_WARPS 16    // 8 -> 16: fwd 6.32 -> 6.22 ms, gE_id 3.4 -> 2.2 ms at 4.4 M items (r02_fusion_notes.txt)
#endif
constexpr int LOADER_WARPS = LGCN_FW_LOADER_WARPS;
constexpr int kThreads = (LOADER_WARPS + 1 + 4) * 32;   // loaders, MMA issuer, epilogue

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes (4 tf32) contiguous; SBO between 8-row
// groups, LBO between the two 16-byte K columns of one K=8 instruction.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// D fp32 (bits 4-5 = 1), A/B tf32 (format 2 at bits 7-9 / 10-12), K-major, N>>3 at 17, M>>4 at 24
// K-major, SWIZZLE_128B: a row is 128 contiguous bytes (32 tf32), its 16-byte unit q stored at
// position q ^ (row & 7); 8-row groups are 1024 bytes apart (SBO), the leading offset is unused.
// The buffer must be 1024-byte aligned (the XOR acts on absolute shared-memory address bits); a
// K = 8 instruction advances the start address by 32 bytes inside the swizzle atom.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((1024u >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
}

template <int D>
struct Smem {
    // per stage: [hi | lo] for A (128 x 32) and B (D x 32), K-major SWIZZLE_128B (128-byte rows)
    float A[NSTAGE][2][MT * KC];
    float B[NSTAGE][2][D * KC];
    unsigned long long full[NSTAGE], empty[NSTAGE], tfull[2], tempty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ void split_store(float *hi_dst, float *lo_dst, const float4 &x) {
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); l.x = x.x - h.x;
    h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); l.y = x.y - h.y;
    h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); l.z = x.z - h.z;
    h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); l.w = x.w - h.w;
    *reinterpret_cast<float4 *>(hi_dst) = h;
    *reinterpret_cast<float4 *>(lo_dst) = l;
}

// EID = true turns the same pipeline into the backward of the item-id block
// (models/lightgcn_fusion.py:45-49 under autograd):
//     gEid[i, k] = sum_o gpre[i, o] W[o, k]  (k < d),   gpre = gH * leaky'(H)
// A rows are gpre rows computed on load (`Eid` := gH, `Cm` := H, K = o: d / 32 chunks); B row k is
// column k of W (the loader transposes by its load pattern: a lane owns one k and reads it for 4
// consecutive o, every LDG.32 a coalesced 128-byte piece of a W row); the epilogue stores the
// accumulator as it is.
template <int D, bool EID>
__global__ void __launch_bounds__(kThreads, 1)
fusion_fwd_tc_kernel(const float *__restrict__ Eid, const float *__restrict__ Cm,
                     const float *__restrict__ W, const float *__restrict__ bias, int64_t n_items,
                     int c, float *__restrict__ H) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem<D> &sm = *reinterpret_cast<Smem<D> *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kin = D + c;
    const int n_chunks = EID ? D / KC : kin / KC;
    const int64_t n_tiles = (n_items + MT - 1) / MT;
    if (smem_u32(smem_raw) & 1023u) __trap();       // SWIZZLE_128B operands need 1024-byte alignment
    constexpr uint32_t IDESC = make_idesc(D);
    constexpr int TMEM_COLS = 2 * D < 32 ? 32 : 2 * D;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(smem_u32(&sm.full[s]), LOADER_WARPS);
            mbar_init(smem_u32(&sm.empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&sm.tfull[b]), 1); mbar_init(smem_u32(&sm.tempty[b]), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == LOADER_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&sm.tmem_base)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp < LOADER_WARPS) {
        // ===== loaders: fp32 rows -> hi/lo tf32 in the SWIZZLE_128B K-major layout =====
        // A chunk is 32 floats = 128 bytes of a row = one swizzle atom row.  Lane (rho = lane / 8,
        // kap = lane % 8) reads the 16-byte piece kap of rows rho and rho + 4 of each 8-row group:
        // every LDG.128 covers 4 rows x 128 contiguous bytes (four full lines; round 1 read 8 rows x
        // 64 bytes per instruction and ran the LSU data pipe at 96 %).  A 128-bit shared store is
        // served per QUARTER warp, so the 8 lanes of one row must hit 32 distinct banks: with the
        // no-swizzle canonical layout (a row's pieces 2 KB apart) they hit the same 4 banks (ncu r02:
        // 396 M conflict wavefronts); in the swizzled layout they write one permuted 128-byte row.
        const int rho = lane >> 3, kap = lane & 7;
        constexpr int NA = (MT / 8) * 2 / LOADER_WARPS;          // A half-groups (4 rows) per warp
        constexpr int NBT = EID ? (D / 32) * (KC / 4) / LOADER_WARPS  // EID: (k block, o quad) tasks
                                : (D / 8) * 2 / LOADER_WARPS;         // B half-groups per warp
        static_assert(NA >= 1 && NBT >= 1, "too many loader warps");
        auto a_row = [&](int j) { const int t = warp * NA + j; return (t >> 1) * 8 + (t & 1) * 4 + rho; };
        auto b_row = [&](int j) { const int t = warp * NBT + j; return (t >> 1) * 8 + (t & 1) * 4 + rho; };
        auto load_chunk = [&](int64_t tile, int ch, float4 (&xa)[NA], float4 (&xb)[NBT]) {
            const int64_t i0 = tile * MT;
            const int k = ch * KC + kap * 4;                     // first float of this lane's piece
#pragma unroll
            for (int j = 0; j < NA; ++j) {
                const int64_t item = i0 + a_row(j);
                xa[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (item < n_items) {
                    if constexpr (EID) {
                        float4 g = ld_stream_f4(Eid + (size_t)item * D + k);
                        const float4 h = ld_stream_f4(Cm + (size_t)item * D + k);
                        g.x *= h.x > 0.f ? 1.f : 0.01f; g.y *= h.y > 0.f ? 1.f : 0.01f;
                        g.z *= h.z > 0.f ? 1.f : 0.01f; g.w *= h.w > 0.f ? 1.f : 0.01f;
                        xa[j] = g;
                    } else {
                        xa[j] = (k < D) ? ld_nc_f4(Eid + (size_t)item * D + k)
                                        : ld_stream_f4(Cm + (size_t)item * c + (k - D));
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NBT; ++j) {
                if constexpr (EID) {
                    // task t = warp * NBT + j: k block t % (D / 32), o quad t / (D / 32); lane = k in block
                    const int t = warp * NBT + j;
                    const float *src = W + (size_t)(ch * KC + (t / (D / 32)) * 4) * kin + (t % (D / 32)) * 32 + lane;
                    xb[j] = make_float4(__ldg(src), __ldg(src + kin), __ldg(src + 2 * kin), __ldg(src + 3 * kin));
                } else {
                    xb[j] = ld_nc_f4(W + (size_t)b_row(j) * kin + k);
                }
            }
        };
        auto store_piece = [&](float *hi_dst, float *lo_dst, int r, const float4 &x) {
            const int off = r * KC + ((kap ^ (r & 7)) << 2);      // floats
            split_store(hi_dst + off, lo_dst + off, x);
        };
        // Three register buffers used IN PLACE: a buffer is refilled with the chunk three ahead right
        // after it has been stored (no register copies that wait for the newest load).  (History,
        // profiles/r02_fusion_notes.txt: with round 1's load mapping -- 8 rows x 64 bytes per LDG.128 --
        // the LSU data pipe was 96 % busy and this rotation measured neutral; after the quarter-warp
        // remapping and the swizzled stores the loaders are latency bound, hence 16 of them.)
        float4 xa[NA], xb[NBT], ya[NA], yb[NBT], za[NA], zb[NBT];
        auto advance = [&](int64_t &t, int &c_) {
            if (++c_ == n_chunks) { c_ = 0; t += gridDim.x; }
        };
        int64_t tile = blockIdx.x, tload = blockIdx.x;            // store cursor / load cursor
        int ch = 0, cload = 0;
        auto load_next = [&](float4 (&a)[NA], float4 (&b)[NBT]) {
            if (tload < n_tiles) load_chunk(tload, cload, a, b);
            advance(tload, cload);
        };
        load_next(xa, xb);
        load_next(ya, yb);
        load_next(za, zb);
        uint32_t it = 0;
        auto step = [&](float4 (&a)[NA], float4 (&b)[NBT]) {
            const int s = it % NSTAGE;
            mbar_wait(smem_u32(&sm.empty[s]), ((it / NSTAGE) & 1) ^ 1);
#pragma unroll
            for (int j = 0; j < NA; ++j) store_piece(&sm.A[s][0][0], &sm.A[s][1][0], a_row(j), a[j]);
#pragma unroll
            for (int j = 0; j < NBT; ++j) {
                if constexpr (EID) {
                    // 8 consecutive lanes = 8 consecutive rows, one unit each at 8 distinct positions
                    const int t = warp * NBT + j;
                    const int r = (t % (D / 32)) * 32 + lane, kq = t / (D / 32);
                    const int off = r * KC + ((kq ^ (r & 7)) << 2);
                    split_store(&sm.B[s][0][off], &sm.B[s][1][off], b[j]);
                } else {
                    store_piece(&sm.B[s][0][0], &sm.B[s][1][0], b_row(j), b[j]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sm.full[s]));
            load_next(a, b);                                  // refill in place: the chunk 3 ahead
            advance(tile, ch);
            ++it;
        };
        while (tile < n_tiles) {
            step(xa, xb);
            if (tile >= n_tiles) break;
            step(ya, yb);
            if (tile >= n_tiles) break;
            step(za, zb);
        }
    } else if (warp == LOADER_WARPS) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t it = 0, tl = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
                const int acc = tl & 1;
                mbar_wait(smem_u32(&sm.tempty[acc]), ((tl >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int ch = 0; ch < n_chunks; ++ch, ++it) {
                    const int s = it % NSTAGE;
                    mbar_wait(smem_u32(&sm.full[s]), (it / NSTAGE) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(&sm.A[s][0][0]), a_lo = smem_u32(&sm.A[s][1][0]);
                    const uint32_t b_hi = smem_u32(&sm.B[s][0][0]), b_lo = smem_u32(&sm.B[s][1][0]);
#pragma unroll
                    for (int kk = 0; kk < KC / 8; ++kk) {
                        const uint64_t ah = make_smem_desc_sw128(a_hi + kk * 32);
                        const uint64_t al = make_smem_desc_sw128(a_lo + kk * 32);
                        const uint64_t bh = make_smem_desc_sw128(b_hi + kk * 32);
                        const uint64_t bl = make_smem_desc_sw128(b_lo + kk * 32);
                        const uint32_t d_addr = tmem_base + acc * D;
                        tc_mma_tf32(d_addr, al, bh, IDESC, (ch > 0 || kk > 0) ? 1u : 0u);   // small terms first
                        tc_mma_tf32(d_addr, ah, bl, IDESC, 1u);
                        tc_mma_tf32(d_addr, ah, bh, IDESC, 1u);
                    }
                    tc_commit(smem_u32(&sm.empty[s]));
                }
                tc_commit(smem_u32(&sm.tfull[acc]));
            }
        }
    } else {
        // ===== epilogue: thread <-> item <-> TMEM lane =====
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        uint32_t tl = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
            const int acc = tl & 1;
            mbar_wait(smem_u32(&sm.tfull[acc]), (tl >> 1) & 1);
            tc_fence_after();
            const int64_t item = tile * MT + r;
#pragma unroll 1
            for (int cb = 0; cb < D / 32; ++cb) {
                float v[32];
                tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * D + cb * 32), v);
                if (item < n_items) {
                    float *dst = H + (size_t)item * D + cb * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        if constexpr (!EID) {
                            o.x += __ldg(bias + cb * 32 + i);
                            o.y += __ldg(bias + cb * 32 + i + 1);
                            o.z += __ldg(bias + cb * 32 + i + 2);
                            o.w += __ldg(bias + cb * 32 + i + 3);
                            o.x = o.x > 0.f ? o.x : 0.01f * o.x;
                            o.y = o.y > 0.f ? o.y : 0.01f * o.y;
                            o.z = o.z > 0.f ? o.z : 0.01f * o.z;
                            o.w = o.w > 0.f ? o.w : 0.01f * o.w;
                        }
                        st_f4(dst + i, o);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sm.tempty[acc]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == LOADER_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

template <int D, bool EID>
int launch_fwd(const float *Eid, const float *C, const float *W, const float *b, int64_t n_items, int c,
               float *H, cudaStream_t st) {
    LGCN_OPT_IN_SMEM((fusion_fwd_tc_kernel<D, EID>), sizeof(Smem<D>));
    const int64_t tiles = (n_items + MT - 1) / MT;
    const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
    fusion_fwd_tc_kernel<D, EID><<<grid, kThreads, sizeof(Smem<D>), st>>>(Eid, C, W, b, n_items, c, H);
    LGCN_LAUNCH_CHECK();
    return 0;
}


// =================================================================================================
// Backward of the projection weights on the tensor cores:
//     gW[o, k] = sum_i gpre[i, o] X[i, k],   gb[o] = sum_i gpre[i, o],   gpre = gH * leaky'(H)
// The reduction runs over the ITEMS:
//     A := X slice  (M = 128 features k0..k0+127 of [E_id | C],  K = items)
//     B := gpre     (N = D output features,                       K = items)
// both K-major in the same canonical no-swizzle layout as the forward kernel (4 consecutive ITEMS
// of one feature are the 16-byte unit).  The tables are item-major in HBM, so a loader lane owns
// one feature and reads it for 4 consecutive items with 4 scalar loads -- every load instruction
// is one coalesced 128-byte row segment (32 features) -- and stores one float4; 8 consecutive
// lanes fill one 128-byte core matrix (conflict free).  (MN-major operand descriptors, which
// would take the float4 loads as they come, returned all-zero accumulators for kind::tf32 with
// the no-swizzle layout on this part, so the transposition is done by the load pattern.)
// Same 3xTF32 split, ring and barriers as the forward kernel.  Grid = (ceil((d+c)/128) feature
// tiles) x (item slabs); a CTA accumulates its slab in TMEM in sub-slabs of kSubChunks chunks
// (two accumulators: the tensor-core fp32 accumulation chain is kept short, partial sums are
// combined by fp32 atomics); epilogue thread <-> TMEM lane <-> feature k.
// =================================================================================================
constexpr int kSubChunks = 32;      // 1024 items (128 accumulations) per TMEM accumulation chain
#ifndef LGCN_BW_LOADER_WARPS
#define LGCN_BW_LOADER_WARPS 16     // the gW loaders are latency bound (ncu r02: long-scoreboard 3.7 per issue)
#endif
constexpr int BW_LOADERS = LGCN_BW_LOADER_WARPS;
constexpr int kBwThreads = (BW_LOADERS + 1 + 4) * 32;   // loaders, MMA issuer, epilogue

template <int D>
struct SmemBw {
    float A[NSTAGE][2][MT * KC];       // X slice, [hi | lo], canonical K-major (K = items)
    float B[NSTAGE][2][D * KC];        // gpre
    unsigned long long full[NSTAGE], empty[NSTAGE], tfull[2], tempty[2];
    uint32_t tmem_base;
};

template <int D>
__global__ void __launch_bounds__(kBwThreads, 1)
fusion_bwd_w_tc_kernel(const float *__restrict__ Eid, const float *__restrict__ Cm,
                       const float *__restrict__ H, const float *__restrict__ gH, int64_t n_items,
                       int c, int64_t items_per_slab, float *__restrict__ gW, float *__restrict__ gb) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemBw<D> &sm = *reinterpret_cast<SmemBw<D> *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kin = D + c;
    const int k0 = blockIdx.x * MT;                              // first X feature of this CTA
    const int64_t ibeg = (int64_t)blockIdx.y * items_per_slab;
    const int64_t iend = min(n_items, ibeg + items_per_slab);
    const int n_chunks = iend > ibeg ? (int)((iend - ibeg + KC - 1) / KC) : 0;
    constexpr uint32_t SBO = 128;
    constexpr uint32_t LBO_A = (MT / 8) * 128;                   // bytes between 4-item K columns of A
    constexpr uint32_t LBO_B = (D / 8) * 128;
    constexpr uint32_t IDESC = make_idesc(D);
    constexpr int TMEM_COLS = 4 * D;                             // {main, lo-terms} x 2 buffers

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(smem_u32(&sm.full[s]), BW_LOADERS);
            mbar_init(smem_u32(&sm.empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&sm.tfull[b]), 1); mbar_init(smem_u32(&sm.tempty[b]), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == BW_LOADERS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&sm.tmem_base)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp < BW_LOADERS) {
        // ===== loaders: a warp task = 32 features (one per lane) x 4 consecutive items =====
        constexpr int NAT = 32 / BW_LOADERS;                     // A tasks per warp: 4 feature blocks x 8 item quads
        constexpr int NBT = (D / 32) * (KC / 4) / BW_LOADERS;    // B tasks per warp
        static_assert(NAT >= 1 && NBT >= 1, "too many loader warps");
        float bsum[NBT];
#pragma unroll
        for (int j = 0; j < NBT; ++j) bsum[j] = 0.f;
        // task t of a warp: feature block (t % NFB), item quad (t / NFB); a lane's feature is fixed
        auto load_chunk = [&](int ch, float4 (&xa)[NAT], float4 (&xg)[NBT], float4 (&xh)[NBT]) {
            const int64_t i0 = ibeg + (int64_t)ch * KC;
#pragma unroll
            for (int j = 0; j < NAT; ++j) {
                const int task = warp * NAT + j;                 // 32 tasks
                const int k = k0 + (task & 3) * 32 + lane;
                const int64_t item = i0 + (task >> 2) * 4;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (k < kin) {
                    const float *src = (k < D) ? Eid + (size_t)item * D + k : Cm + (size_t)item * c + (k - D);
                    const size_t ld = (k < D) ? (size_t)D : (size_t)c;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (item + q < iend) v[q] = __ldg(src + q * ld);
                }
                xa[j] = make_float4(v[0], v[1], v[2], v[3]);
            }
#pragma unroll
            for (int j = 0; j < NBT; ++j) {
                const int task = warp * NBT + j;
                const int o = (task % (D / 32)) * 32 + lane;
                const int64_t item = i0 + (task / (D / 32)) * 4;
                float g[4] = {0.f, 0.f, 0.f, 0.f}, h[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (item + q < iend) {
                        g[q] = __ldg(gH + (size_t)(item + q) * D + o);
                        h[q] = __ldg(H + (size_t)(item + q) * D + o);
                    }
                xg[j] = make_float4(g[0], g[1], g[2], g[3]);
                xh[j] = make_float4(h[0], h[1], h[2], h[3]);
            }
        };
        // Software pipeline: x = chunk ch, finished (gpre formed); r = the RAW loads of chunk ch + 1,
        // issued one whole iteration earlier.  Per iteration: finish r into y (the only wait, for loads
        // that have had a full store phase to arrive), re-issue r for chunk ch + 2, store x, x = y.
        // (Round 2 history, 4.4 M items, whole backward: one-ahead loads waited for at the end of
        // their own iteration 13.5 ms with 8 loader warps, 12.15 with 16; an in-place two-buffer
        // rotation was slower, 20.1 vs 16.9 -- profiles/r02_fusion_notes.txt.)
        float4 xa[NAT], xg[NBT], ra[NAT], rg[NBT], rh[NBT];
        auto finish = [&](float4 (&da)[NAT], float4 (&dg)[NBT]) {
#pragma unroll
            for (int j = 0; j < NAT; ++j) da[j] = ra[j];
#pragma unroll
            for (int j = 0; j < NBT; ++j) {
                float4 g = rg[j];
                g.x *= rh[j].x > 0.f ? 1.f : 0.01f; g.y *= rh[j].y > 0.f ? 1.f : 0.01f;
                g.z *= rh[j].z > 0.f ? 1.f : 0.01f; g.w *= rh[j].w > 0.f ? 1.f : 0.01f;
                dg[j] = g;
            }
        };
        if (0 < n_chunks) { load_chunk(0, ra, rg, rh); finish(xa, xg); }
        if (1 < n_chunks) load_chunk(1, ra, rg, rh);
        for (int ch = 0; ch < n_chunks; ++ch) {
            float4 ya[NAT], yg[NBT];
            if (ch + 1 < n_chunks) finish(ya, yg);
            if (ch + 2 < n_chunks) load_chunk(ch + 2, ra, rg, rh);   // in flight for a whole iteration
            const int s = ch % NSTAGE;
            mbar_wait(smem_u32(&sm.empty[s]), ((ch / NSTAGE) & 1) ^ 1);
#pragma unroll
            for (int j = 0; j < NAT; ++j) {
                const int task = warp * NAT + j;
                const int r = (task & 3) * 32 + lane, kq = task >> 2;
                const int off = kq * (MT / 8) * 32 + (r >> 3) * 32 + (r & 7) * 4;       // floats
                split_store(&sm.A[s][0][off], &sm.A[s][1][off], xa[j]);
            }
#pragma unroll
            for (int j = 0; j < NBT; ++j) {
                const int task = warp * NBT + j;
                const float4 g = xg[j];
                bsum[j] += (g.x + g.y) + (g.z + g.w);
                const int r = (task % (D / 32)) * 32 + lane, kq = task / (D / 32);
                const int off = kq * (D / 8) * 32 + (r >> 3) * 32 + (r & 7) * 4;
                split_store(&sm.B[s][0][off], &sm.B[s][1][off], g);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sm.full[s]));
#pragma unroll
            for (int j = 0; j < NAT; ++j) xa[j] = ya[j];
#pragma unroll
            for (int j = 0; j < NBT; ++j) xg[j] = yg[j];
        }
        // bias gradient: once per item slab (feature tile 0 only)
        if (blockIdx.x == 0) {
#pragma unroll
            for (int j = 0; j < NBT; ++j)
                atomicAdd(gb + ((warp * NBT + j) % (D / 32)) * 32 + lane, bsum[j]);
        }
    } else if (warp == BW_LOADERS) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int sub = 0;
            for (int ch = 0; ch < n_chunks; ++ch) {
                const int acc = sub & 1;
                if (ch % kSubChunks == 0) {
                    mbar_wait(smem_u32(&sm.tempty[acc]), ((sub >> 1) & 1) ^ 1);
                    tc_fence_after();
                }
                const int s = ch % NSTAGE;
                mbar_wait(smem_u32(&sm.full[s]), (ch / NSTAGE) & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(&sm.A[s][0][0]), a_lo = smem_u32(&sm.A[s][1][0]);
                const uint32_t b_hi = smem_u32(&sm.B[s][0][0]), b_lo = smem_u32(&sm.B[s][1][0]);
#pragma unroll
                for (int kk = 0; kk < KC / 8; ++kk) {                     // 8 items per MMA
                    const uint64_t ah = make_smem_desc(a_hi + kk * 2 * LBO_A, LBO_A, SBO);
                    const uint64_t al = make_smem_desc(a_lo + kk * 2 * LBO_A, LBO_A, SBO);
                    const uint64_t bh = make_smem_desc(b_hi + kk * 2 * LBO_B, LBO_B, SBO);
                    const uint64_t bl = make_smem_desc(b_lo + kk * 2 * LBO_B, LBO_B, SBO);
                    // the two small cross terms go to their own accumulator: the tensor core
                    // truncates its fp32 accumulator on every instruction, so the bias grows with
                    // the chain length (measured 5e-5 over 3072 accumulations); this keeps the
                    // main chain at one accumulation per 8 items
                    const uint32_t d_main = tmem_base + acc * D, d_lo = tmem_base + (2 + acc) * D;
                    const uint32_t first = (ch % kSubChunks > 0 || kk > 0) ? 1u : 0u;
                    tc_mma_tf32(d_lo, al, bh, IDESC, first);
                    tc_mma_tf32(d_lo, ah, bl, IDESC, 1u);
                    tc_mma_tf32(d_main, ah, bh, IDESC, first);
                }
                tc_commit(smem_u32(&sm.empty[s]));
                if (ch % kSubChunks == kSubChunks - 1 || ch == n_chunks - 1) {
                    tc_commit(smem_u32(&sm.tfull[acc]));
                    ++sub;
                }
            }
        }
    } else {
        // ===== epilogue: thread <-> TMEM lane <-> X feature k; columns = output features o =====
        const int quad = warp & 3;
        const int k = k0 + quad * 32 + lane;
        const int n_sub = (n_chunks + kSubChunks - 1) / kSubChunks;
        for (int sub = 0; sub < n_sub; ++sub) {
            const int acc = sub & 1;
            mbar_wait(smem_u32(&sm.tfull[acc]), (sub >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cb = 0; cb < D / 32; ++cb) {
                float v[32], w[32];
                tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * D + cb * 32), v);
                tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((2 + acc) * D + cb * 32), w);
                if (k < kin) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)                  // a warp adds 32 consecutive k: coalesced
                        atomicAdd(gW + (size_t)(cb * 32 + i) * kin + k, v[i] + w[i]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sm.tempty[acc]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == BW_LOADERS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

template <int D>
int launch_bwd_w(const float *Eid, const float *C, const float *H, const float *gH, int64_t n_items,
                 int c, float *gW, float *gb, cudaStream_t st) {
    LGCN_OPT_IN_SMEM((fusion_bwd_w_tc_kernel<D>), sizeof(SmemBw<D>));
    const int ktiles = (D + c + MT - 1) / MT;
    int64_t slabs = kNumSMs / ktiles;                            // one wave of CTAs (1 CTA per SM)
    if (slabs < 1) slabs = 1;
    int64_t per = (n_items + slabs - 1) / slabs;
    per = (per + KC - 1) / KC * KC;
    slabs = (n_items + per - 1) / per;
    dim3 grid((unsigned)ktiles, (unsigned)slabs);
    fusion_bwd_w_tc_kernel<D><<<grid, kBwThreads, sizeof(SmemBw<D>), st>>>(Eid, C, H, gH, n_items, c, per, gW, gb);
    LGCN_LAUNCH_CHECK();
    return 0;
}

}  // namespace ftc
}  // namespace lgcn

// used by lgcn_fusion.cu: returns -100 when the shape is not handled by the tensor-core path
extern "C" __attribute__((visibility("hidden"))) int lgcn_fusion_fwd_tc_try(
    const float *Eid, const float *C, const float *W, const float *b, int64_t n_items, int32_t d,
    int32_t c, float *H, cudaStream_t st) {
    using namespace lgcn::ftc;
    if ((d != 64 && d != 128) || c % KC != 0 || n_items < MT) return -100;
    return d == 64 ? launch_fwd<64, false>(Eid, C, W, b, n_items, c, H, st)
                   : launch_fwd<128, false>(Eid, C, W, b, n_items, c, H, st);
}

// gEid of the backward pass: the forward pipeline with A = gH * leaky'(H) and B = W[:, :d]^T
extern "C" __attribute__((visibility("hidden"))) int lgcn_fusion_bwd_eid_tc_try(
    const float *W, const float *H, const float *gH, int64_t n_items, int32_t d, int32_t c, float *gEid,
    cudaStream_t st) {
    using namespace lgcn::ftc;
    if ((d != 64 && d != 128) || n_items < MT) return -100;
    return d == 64 ? launch_fwd<64, true>(gH, H, W, nullptr, n_items, c, gEid, st)
                   : launch_fwd<128, true>(gH, H, W, nullptr, n_items, c, gEid, st);
}

// gW / gb of the backward pass (accumulated into the caller's zero-initialised or running buffers)
extern "C" __attribute__((visibility("hidden"))) int lgcn_fusion_bwd_w_tc_try(
    const float *Eid, const float *C, const float *H, const float *gH, int64_t n_items, int32_t d,
    int32_t c, float *gW, float *gb, cudaStream_t st) {
    using namespace lgcn::ftc;
    if ((d != 64 && d != 128) || c % 4 != 0 || n_items < MT) return -100;
    return d == 64 ? launch_bwd_w<64>(Eid, C, H, gH, n_items, c, gW, gb, st)
                   : launch_bwd_w<128>(Eid, C, H, gH, n_items, c, gW, gb, st);
}
