// lgcn_spmm_impl.cuh -- normalised-adjacency CSR SpMM with fused epilogues (sm_100a): the kernels.
// Compiled once per table width by lgcn_spmm_d{16,32,64,128,256}.cu (parallel translation units);
// lgcn_spmm.cu holds the C entry points.
//
// Replaces torch.sparse.mm (reference models/lightgcn.py:45, models/lightgcn_fusion.py:56),
// its autograd backward (A is symmetric, so the same kernel serves), the layer mean
// (reference models/lightgcn.py:54) and, in ADAM mode, optimizer.step() (reference
// main.py:526).  HBM-bound gather/stream work: no tensor cores.
//
// Mapping.  A table row of d floats is owned by a "worker" = sub-warp group of d/4 lanes
// (float4 per lane; one warp at d=128, a half warp at d=64).  Every lane owns fixed feature
// columns, so one output row is a SEQUENTIAL fp32 FMA chain in ascending column order --
// bit-equal to the CPU reference.
//
// A worker owns a CHUNK of R consecutive rows, whose entries are one contiguous range of the
// packed {col,val} array.  It walks that range as a flat stream: {col,val} tiles are loaded
// coalesced one tile ahead, the gathers of X rows are issued UNROLL deep ahead of the FMAs and
// do not drain at row boundaries (average degree is ~4, so per-row pipelines would be latency
// bound), and the row an entry belongs to is found with one ballot over the per-lane row ends.
// Finished rows are staged in shared memory; the epilogue then streams the chunk's R rows with
// all of its operand loads independent (mean of the earlier layers / Horner addend / Adam).
#pragma once
#include <limits.h>
#include <stdio.h>

#include <type_traits>

#include "lgcn_common.cuh"

namespace lgcn {

constexpr int kUnroll = 8;
#ifndef LGCN_SPMM_WARPS
#define LGCN_SPMM_WARPS 4
#endif
#ifndef LGCN_SPMM_MINBLOCKS
#define LGCN_SPMM_MINBLOCKS 1          // MEAN / ADAM epilogues need the registers
#endif
#ifndef LGCN_SPMM_MINBLOCKS_LIGHT
#define LGCN_SPMM_MINBLOCKS_LIGHT 1    // (8 = cap at 64 registers: measured slower, 7.02 vs 6.47 ms)
#endif
#ifndef LGCN_SPMM_MINBLOCKS_SMALL_MEAN
#define LGCN_SPMM_MINBLOCKS_SMALL_MEAN 6    // (the MEAN epilogue spills at 64 registers)
#endif
#ifndef LGCN_SPMM_MINBLOCKS_SMALL
#define LGCN_SPMM_MINBLOCKS_SMALL 8    // small (latency-bound) graphs want the warps: 64 registers,
#endif                                 // Gowalla step 0.615 -> 0.565 ms (6 blocks: 0.586)
constexpr int kWarps = LGCN_SPMM_WARPS;
constexpr int kThreads = kWarps * 32;

template <int D, int RSEL>
struct ChunkCfg {
    using G = RowGeom<D>;
    // rows per worker chunk: RSEL == 0 -> up to 16 rows (big graphs: amortise the per-chunk
    // pointer loads), RSEL == 1 -> 4-row chunks (small graphs: more workers, shorter chains).
    // One row end per lane (R <= LANES).  Several row ends per lane (R > LANES for narrow
    // tables) was measured slower at d=16/32 and cost 8 registers at d=128.
    static constexpr int RMAX = (2048 / D) < 4 ? 4 : (2048 / D);
    static constexpr int RBIG = G::LANES < RMAX ? G::LANES : RMAX;
    static constexpr int R = RSEL == 0 ? RBIG : 4;
    static constexpr int WORKERS = kWarps * G::GROUPS;                // workers per CTA
    static constexpr int ROWS_PER_CTA = WORKERS * R;
    static constexpr int UMAX = G::VEC > 1 ? 4 : 8;                   // gathers per batch
    static constexpr int U = G::LANES < UMAX ? G::LANES : UMAX;
    static constexpr size_t SMEM = (size_t)ROWS_PER_CTA * D * sizeof(float);
};

// loads / stores of streamed (touched once) data, with or without the L2 evict_first hint
template <bool HINT>
__device__ __forceinline__ float4 ld_s(const float *p, uint64_t pol) {
    return HINT ? ld_stream_f4_hint(p, pol) : ld_stream_f4(p);
}
template <bool HINT>
__device__ __forceinline__ void st_s(float *p, const float4 &v, uint64_t pol) {
    if (HINT) st_f4_hint(p, v, pol); else st_f4(p, v);
}
template <bool HINT>
__device__ __forceinline__ int2 ld_cv(const int2 *p, uint64_t pol) {
    return HINT ? ld_stream_i2_hint(p, pol) : __ldg(p);
}

// Column classes of the graph plan (include/lgcn.h: bits 31/30 of lgcn_colval.col).  A gathered
// row of a HOT column (one of the highest-degree nodes, as many as fit the L2 budget) is kept with
// evict_last, the row of a column referenced once in the whole launch leaves L2 first, the rest
// is evict_normal (or evict_first with LGCN_SPMM_F_COLD_FIRST: reuse distances of mid-degree
// columns are far beyond L2 at the Amazon shape).
struct GatherPolicy {
    uint64_t last, mid, first;
};
template <bool HINT>
__device__ __forceinline__ GatherPolicy gather_policy(int flags, uint64_t pol_first) {
    GatherPolicy g{0ull, 0ull, 0ull};
    if (HINT) {
        g.first = pol_first;
        g.last = policy_evict_last();
        g.mid = (flags & LGCN_SPMM_F_COLD_FIRST) ? pol_first : policy_evict_normal();
    }
    return g;
}
__device__ __forceinline__ uint64_t pick_policy(const GatherPolicy &g, int col_raw) {
    return col_raw < 0 ? g.last : ((col_raw & LGCN_COL_ONCE) ? g.first : g.mid);
}

// ---- layer-0 override / ADAM row skip (lgcn_spmm_args.x_alt, g_skip) --------------------------
// Both ranges are tested with one unsigned compare; the alternative base pointer is pre-shifted by
// the range start so that the SAME row index addresses either table.
struct AltRange {
    uint32_t lo, n;                 // rows [lo, lo + n); n == 0: never taken
#ifdef LGCN_SPMM_NO_ALT     // A/B only (profiles/build_variant.sh): what the override costs when unused
    __device__ __forceinline__ bool has(int) const { return false; }
#else
    __device__ __forceinline__ bool has(int r) const { return (uint32_t)r - lo < n; }
#endif
};
// ALT = false instantiations (every LightGCN call) compile the override out: measured at the Amazon
// shape the run-time test alone cost 1.5 % of a PLAIN and 9.5 % of a MEAN launch
// (profiles/r02_alt_ab.txt).
template <bool ALT>
__device__ __forceinline__ AltRange alt_x_range(const lgcn_spmm_args &a) {
    if (!ALT) return AltRange{0u, 0u};
    return AltRange{(uint32_t)a.alt_begin, (a.flags & LGCN_SPMM_F_ALT_X) ? (uint32_t)a.alt_rows : 0u};
}
template <bool ALT>
__device__ __forceinline__ AltRange alt_layer0_range(const lgcn_spmm_args &a) {
    if (!ALT) return AltRange{0u, 0u};
    return AltRange{(uint32_t)a.alt_begin, (a.flags & LGCN_SPMM_F_ALT_LAYER0) ? (uint32_t)a.alt_rows : 0u};
}
template <bool ALT>      // ADAM instantiations: ALT = "the call carries g_skip" (2 % of a plain ADAM hop)
__device__ __forceinline__ AltRange adam_skip_range(const lgcn_spmm_args &a) {
    if (!ALT) return AltRange{0u, 0u};
    return AltRange{(uint32_t)a.skip_begin, a.g_skip ? (uint32_t)a.skip_rows : 0u};
}
template <int D>
__device__ __forceinline__ const float *alt_shifted(const lgcn_spmm_args &a) {      // x_alt - alt_begin rows
    return a.x_alt ? a.x_alt - (size_t)a.alt_begin * D : a.x;
}

// ---- L2 prefetch of the epilogue's operand rows ---------------------------------------------
// A worker first walks its chunk's entries (gather phase), then streams the chunk's rows of the
// epilogue operands (p/m/v for ADAM, the earlier layers for MEAN, a dense addend).  Both phases
// are latency bound inside the warp and only overlap ACROSS warps; at 20-25 warps per SM the
// epilogue's first-touch DRAM loads left the ADAM hop at 66 % DRAM utilisation (ncu r01).  The rows
// of a chunk are contiguous, so ONE bulk prefetch per operand table, issued while the last
// entries of the chunk are still being gathered, turns those loads into L2 hits ~1-2 us later
// (footprint in L2: bandwidth x lead time ~ 10 MB).  Only when the tables stream from HBM (HINT);
// disabled by LGCN_SPMM_F_NO_PREFETCH.
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}
template <int D, int MODE>
__device__ __forceinline__ void prefetch_epilogue_rows(const lgcn_spmm_args &a, int64_t r0, int nvr) {
    if (nvr <= 0 || (a.flags & LGCN_SPMM_F_NO_PREFETCH)) return;
    const size_t off = (size_t)r0 * D;
    const uint32_t bytes = (uint32_t)nvr * D * 4;
    // Measured at the Amazon shape (profiles/r02_prefetch_ab.txt): MEAN 9.27 -> 8.58 ms, dense ADD
    // 6.30 -> 6.04 ms; ADAM (p, m, v: 24 KB per worker, held for its long epilogue) gains nothing
    // and loses in the engine's step, and tables narrower than d = 64 lose 4 % -- both excluded.
    if (D < 64 || MODE == LGCN_SPMM_ADAM) return;
    if (MODE == LGCN_SPMM_MEAN) {
        for (int l = 0; l < a.n_layers; ++l) prefetch_l2_bulk(a.layers[l] + off, bytes);
    } else if (MODE == LGCN_SPMM_ADD) {
        if (!a.addend_rowflag) prefetch_l2_bulk(a.addend + off, bytes);     // flagged = mostly zero rows
    }
}

// ---- epilogue for one row held in registers (long-row combine path) -----------------------
template <int D, int MODE, bool ALT>
__device__ __forceinline__ void epilogue_row(const lgcn_spmm_args &a, int64_t row,
                                             float4 (&acc)[RowGeom<D>::VEC]) {
    using G = RowGeom<D>;
    const int sub = (threadIdx.x & 31) % G::LANES;
#pragma unroll
    for (int v = 0; v < G::VEC; ++v) {
        const size_t off = (size_t)row * D + sub * 4 + v * G::LANES * 4;
        if (MODE == LGCN_SPMM_PLAIN) {
            st_f4(a.y + off, acc[v]);
        } else if (MODE == LGCN_SPMM_ADD) {
            float4 g = ld_stream_f4(a.addend + off);
            add4(g, acc[v]);
            st_f4(a.y + off, g);
        } else if (MODE == LGCN_SPMM_MEAN) {
            const float *l0 = alt_layer0_range<ALT>(a).has((int)row) ? alt_shifted<D>(a) : a.layers[0];
            float4 s = ld_stream_f4(l0 + off);
            for (int l = 1; l < a.n_layers; ++l) {
                const float4 t = ld_stream_f4(a.layers[l] + off);
                add4(s, t);
            }
            add4(s, acc[v]);
            const float div = (float)(a.n_layers + 1);
            s.x = __fdiv_rn(s.x, div); s.y = __fdiv_rn(s.y, div);
            s.z = __fdiv_rn(s.z, div); s.w = __fdiv_rn(s.w, div);
            st_f4(a.y + off, s);
        } else {  // LGCN_SPMM_ADAM
            float4 g = acc[v];
            if (a.addend) { const float4 t = ld_stream_f4(a.addend + off); add4(g, t); }
            if (adam_skip_range<ALT>(a).has((int)row)) {      // not a parameter row: hand the gradient on
                st_f4(a.g_skip - (size_t)a.skip_begin * D + off, g);
                continue;
            }
            if (a.addend2) { const float4 t = ld_stream_f4(a.addend2 + off); add4(g, t); }
            float4 p = *reinterpret_cast<const float4 *>(a.p + off);
            float4 m = *reinterpret_cast<const float4 *>(a.m + off);
            float4 vv = *reinterpret_cast<const float4 *>(a.v + off);
            const float ss = __ldg(a.adam_scalars), bs = __ldg(a.adam_scalars + 1);
            adam4(p, m, vv, g, ss, bs, a.beta1, a.beta2, a.eps);
            st_f4(a.p + off, p); st_f4(a.m + off, m); st_f4(a.v + off, vv);
            if (a.g_out) st_f4(a.g_out + off, g);
        }
    }
}

// ---- chunk epilogue: stream the staged rows, operand loads batched ahead of the math ------
#ifndef LGCN_MEAN_B4
#define LGCN_MEAN_B4 4                 // rows per MEAN epilogue batch when at most 4 layers are read
#endif
// NLM: compile-time bound of the number of earlier layers the MEAN epilogue reads (8 = the ABI
// limit; 4 covers K <= 4, i.e. every configuration of the reference: half the operand registers,
// spent on twice the rows per batch = twice the loads in flight)
template <int D, int MODE, int R, bool HINT, bool ALT, int NLM = 8>
__device__ __forceinline__ void chunk_epilogue(const lgcn_spmm_args &a, const float *stage,
                                               int64_t r0, int nvr, unsigned long_bits, uint64_t pol,
                                               const unsigned (&rfw)[R / 4], unsigned wmask = 0xffffffffu) {
    using G = RowGeom<D>;
    const int sub = (threadIdx.x & 31) % G::LANES;
    // rows per batch.  ADAM: 2 rows of g / g2 / p / m / v are 40 operand registers; 4 rows lose 23 % at
    // d = 128 (14.18 vs 11.49 ms) but win 7 % at d = 16 (2.27 vs 2.44 ms); d = 32 loses again (4.55 vs 3.45 ms)
    constexpr int B = MODE == LGCN_SPMM_ADAM ? ((D <= 16 && R % 4 == 0) ? 4 : 2)
                                              : (MODE == LGCN_SPMM_MEAN ? (NLM <= 4 ? LGCN_MEAN_B4 : 2) : 4);
    static_assert(R % B == 0, "chunk rows must be a multiple of the epilogue batch");
    const float div = (float)(a.n_layers + 1);
    float ss = 0.f, bs = 1.f;
    if (MODE == LGCN_SPMM_ADAM) { ss = __ldg(a.adam_scalars); bs = __ldg(a.adam_scalars + 1); }
#pragma unroll
    for (int v = 0; v < G::VEC; ++v) {
        const int coff = sub * 4 + v * G::LANES * 4;
        for (int rb = 0; rb < R; rb += B) {
            bool on[B];
            size_t off[B];
#pragma unroll
            for (int i = 0; i < B; ++i) {
                const int rr = rb + i;
                on[i] = rr < nvr && !((long_bits >> rr) & 1u) && ((wmask >> rr) & 1u);
                off[i] = (size_t)(r0 + rr) * D + coff;
            }
            [[maybe_unused]] const AltRange alt0 = alt_layer0_range<ALT>(a), skip = adam_skip_range<ALT>(a);
            if (MODE == LGCN_SPMM_PLAIN) {
#pragma unroll
                for (int i = 0; i < B; ++i)
                    if (on[i]) st_s<HINT>(a.y + off[i], *reinterpret_cast<const float4 *>(stage + (rb + i) * D + coff), pol);
            } else if (MODE == LGCN_SPMM_ADD) {
                float4 t[B];
#pragma unroll
                for (int i = 0; i < B; ++i)
                    if (on[i]) {
                        // all-zero addend rows (no gradient landed there) are not read at all
                        // (redirecting them to one zero row made every SM hammer the same L2
                        // line with no-allocate loads: 7.0 ms vs 6.1 ms dense at the Amazon shape)
                        const bool nz = ((rfw[(rb + i) >> 2] >> (((rb + i) & 3) * 8)) & 0xffu) != 0;
                        t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (nz) t[i] = ld_s<HINT>(a.addend + off[i], pol);
                    }
#pragma unroll
                for (int i = 0; i < B; ++i)
                    if (on[i]) {
                        const float4 y = *reinterpret_cast<const float4 *>(stage + (rb + i) * D + coff);
                        add4(t[i], y);
                        st_s<HINT>(a.y + off[i], t[i], pol);
                    }
            } else if (MODE == LGCN_SPMM_MEAN) {
                float4 t[B][NLM];
#pragma unroll
                for (int i = 0; i < B; ++i)
#pragma unroll
                    for (int l = 0; l < NLM; ++l)
                        if (on[i] && l < a.n_layers) {
                            const float *lp = (l == 0 && alt0.has((int)(r0 + rb + i))) ? alt_shifted<D>(a) : a.layers[l];
                            t[i][l] = ld_s<HINT>(lp + off[i], pol);
                        }
#pragma unroll
                for (int i = 0; i < B; ++i)
                    if (on[i]) {
                        float4 s = t[i][0];
#pragma unroll
                        for (int l = 1; l < NLM; ++l) if (l < a.n_layers) add4(s, t[i][l]);
                        const float4 y = *reinterpret_cast<const float4 *>(stage + (rb + i) * D + coff);
                        add4(s, y);
                        s.x = __fdiv_rn(s.x, div); s.y = __fdiv_rn(s.y, div);
                        s.z = __fdiv_rn(s.z, div); s.w = __fdiv_rn(s.w, div);
                        st_s<HINT>(a.y + off[i], s, pol);
                    }
            } else {  // ADAM
                float4 g[B], p[B], m[B], vv[B], g2[B];
                bool sk[B];
#pragma unroll
                for (int i = 0; i < B; ++i) {
                    sk[i] = skip.has((int)(r0 + rb + i));
                    if (on[i]) {
                        const bool nz = ((rfw[(rb + i) >> 2] >> (((rb + i) & 3) * 8)) & 0xffu) != 0;
                        g[i] = g2[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (a.addend && nz) g[i] = ld_s<HINT>(a.addend + off[i], pol);
                        if (sk[i]) continue;               // not a parameter row: no p / m / v traffic
                        if (a.addend2 && nz) g2[i] = ld_s<HINT>(a.addend2 + off[i], pol);
                        p[i] = ld_s<HINT>(a.p + off[i], pol);
                        m[i] = ld_s<HINT>(a.m + off[i], pol);
                        vv[i] = ld_s<HINT>(a.v + off[i], pol);
                    }
                }
#pragma unroll
                for (int i = 0; i < B; ++i)
                    if (on[i]) {
                        float4 gr = *reinterpret_cast<const float4 *>(stage + (rb + i) * D + coff);
                        if (a.addend) add4(gr, g[i]);
                        if (sk[i]) {                       // hand the gradient to the projection's backward
                            st_s<HINT>(a.g_skip - (size_t)a.skip_begin * D + off[i], gr, pol);
                            continue;
                        }
                        if (a.addend2) add4(gr, g2[i]);
                        adam4(p[i], m[i], vv[i], gr, ss, bs, a.beta1, a.beta2, a.eps);
                        st_s<HINT>(a.p + off[i], p[i], pol);
                        st_s<HINT>(a.m + off[i], m[i], pol);
                        st_s<HINT>(a.v + off[i], vv[i], pol);
                        if (a.g_out) st_s<HINT>(a.g_out + off[i], gr, pol);
                    }
            }
        }
    }
}

// Sparse output of a flagged hop (lgcn_spmm_args.y_rowflag): a row is written only if it summed a
// flagged row of x (bit in `touched`) or has a flagged addend; y_rowflag tells the consumer (the
// next hop's x_rowflag) which rows exist.  Long rows are always written, densely, by the combine
// kernel.  Returns the mask of the chunk's rows to write.
template <int R>
__device__ __forceinline__ unsigned sparse_out_mask(const lgcn_spmm_args &a, const unsigned (&rfw)[R / 4],
                                                    unsigned touched, int64_t r0, int sub, int nvr,
                                                    bool my_long) {
    if (!a.y_rowflag) return 0xffffffffu;
    unsigned wmask = touched;
#pragma unroll
    for (int rr = 0; rr < R; ++rr)
        if ((rfw[rr >> 2] >> ((rr & 3) * 8)) & 0xffu) wmask |= 1u << rr;
    if (sub < nvr) a.y_rowflag[r0 + sub] = (uint8_t)(((wmask >> sub) & 1u) | (my_long ? 1u : 0u));
    return wmask;
}

template <int D, bool ALT, int FMODE = -1>
__device__ __forceinline__ void long_seg_body(const lgcn_spmm_args &a, int64_t block);
template <int D, int MODE, bool ALT, int CBMAX>
__device__ __forceinline__ void combine_long_row(const lgcn_spmm_args &a, int i, int sub);

// ---- main kernel: one worker per chunk of R rows ---------------------------------------------
template <int D, int MODE, int RSEL, bool HINT, bool XF, bool ALT>
__global__ void __launch_bounds__(kThreads, RSEL == 1 ? (MODE == LGCN_SPMM_MEAN ? LGCN_SPMM_MINBLOCKS_SMALL_MEAN : LGCN_SPMM_MINBLOCKS_SMALL) :
                                  (MODE == LGCN_SPMM_PLAIN || MODE == LGCN_SPMM_ADD) ? LGCN_SPMM_MINBLOCKS_LIGHT : LGCN_SPMM_MINBLOCKS)
spmm_chunk_kernel(const __grid_constant__ lgcn_spmm_args a) {
    using G = RowGeom<D>;
    using C = ChunkCfg<D, RSEL>;
    [[maybe_unused]] const int64_t seg_blocks =
        (RSEL == 1 && a.n_long > 0) ? (a.n_seg + kWarps * G::GROUPS - 1) / (kWarps * G::GROUPS) : 0;
    if constexpr (RSEL == 1) {
        // small (L2-resident, latency-bound) graphs: the long-row segment workers ride in the same
        // launch as extra CTAs, so the two independent phases overlap (Gowalla shape: the separate
        // 31 us segment launch was a quarter of an SpMM call)
        // They take the FIRST blocks of the grid: 128-entry segments are the longest work items, and
        // with long_done the rows they finish are combined while the chunk workers still run.
        if ((int64_t)blockIdx.x < seg_blocks) {
            long_seg_body<D, ALT, MODE>(a, (int64_t)blockIdx.x);
            return;
        }
    }
    const uint64_t pol = HINT ? policy_evict_first() : 0ull;
    const GatherPolicy gpol = gather_policy<HINT && !XF>(a.flags, pol);
    extern __shared__ __align__(16) float stage_all[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int grp = lane / G::LANES;
    const int sub = lane % G::LANES;
    const int gshift = grp * G::LANES;
    const unsigned gbits = (G::LANES == 32) ? 0xffffffffu : ((1u << G::LANES) - 1u);
    float *stage = stage_all + (size_t)((warp * G::GROUPS + grp) * C::R) * D;

    const int64_t worker = (((int64_t)blockIdx.x - seg_blocks) * kWarps + warp) * G::GROUPS + grp;
    int64_t chunk = worker;
    if constexpr (RSEL == 1) {      // optional plan: chunks in descending length (equal lengths share a warp)
        if (a.chunk_order && worker * C::R < a.n_rows) chunk = __ldg(a.chunk_order + worker);
    }
    const int64_t r0 = chunk * C::R;
    const int64_t left = a.n_rows - r0;
    const int nvr = left <= 0 ? 0 : (left < C::R ? (int)left : C::R);

    // "row received a gradient" flags of the chunk's addend rows (bytes r0 .. r0+R, the array is
    // padded): fetched up front so that the epilogue has no dependent flag -> addend load chain
    unsigned rfw[C::R / 4];
#pragma unroll
    for (int i = 0; i < C::R / 4; ++i)
        rfw[i] = ((MODE == LGCN_SPMM_ADD || MODE == LGCN_SPMM_ADAM) && a.addend_rowflag && nvr > 0)
                     ? __ldg(reinterpret_cast<const unsigned *>(a.addend_rowflag + r0) + i) : 0xffffffffu;
    unsigned rb = 0, re = 0;
    if (sub < nvr) {
        rb = __ldg(a.rowptr + r0 + sub);
        re = __ldg(a.rowptr + r0 + sub + 1);
    }
    const bool my_long = (rb >> 31) != 0;
    const int my_beg = (int)(rb & 0x7fffffffu);
    int my_end = (int)(re & 0x7fffffffu);
    int chunk_beg = __shfl_sync(0xffffffffu, my_beg, 0, G::LANES);
    int chunk_end = __shfl_sync(0xffffffffu, my_end, nvr > 0 ? nvr - 1 : 0, G::LANES);
    if (nvr == 0) chunk_beg = chunk_end = 0;
    if (sub >= nvr) my_end = INT_MAX;                    // sentinel: never passed
    const unsigned long_bits = (__ballot_sync(0xffffffffu, my_long) >> gshift) & gbits;

    const int n_e = chunk_end - chunk_beg;
    int max_n = n_e;
#pragma unroll
    for (int off = G::LANES; off < 32; off <<= 1)
        max_n = max(max_n, __shfl_xor_sync(0xffffffffu, max_n, off));

    const int2 *cvp = reinterpret_cast<const int2 *>(a.colval) + chunk_beg;
    int2 cv = make_int2(0, 0);
    if (sub < n_e) cv = ld_cv<HINT>(cvp + sub, pol);
    [[maybe_unused]] const AltRange altx = alt_x_range<ALT>(a);
    [[maybe_unused]] const float *xalt = alt_shifted<D>(a);

    float4 acc[G::VEC];
#pragma unroll
    for (int v = 0; v < G::VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = 0;                                          // row (within the chunk) being summed
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    // Small graphs (RSEL == 1, L2-resident tables): the walk is INSTRUCTION bound -- ncu r02 at the
    // Gowalla shape: 65 % issue-active, 47 warp instructions per two-entry step of which 4 FFMA and
    // 1 LDG, the rest the per-entry row bookkeeping (ballot / shift / popc / compares) and address
    // selects.  Here every lane prepares ITS entry of a tile once: `meta` = the chunk row the entry
    // belongs to (the 4 row ends sit in registers; kSkip for slots past the chunk's end) and `goff`
    // = the gathered row's BYTE offset (small tables: < 2^31; bit 31: the row comes from x_alt); the per-entry
    // work is then three shuffles, one compare, the load and the FMAs.
    constexpr bool FAST = RSEL == 1;
    constexpr int kSkip = 7;
    [[maybe_unused]] int ends[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX};
    [[maybe_unused]] int meta = kSkip;
    [[maybe_unused]] unsigned goff = 0u;
    [[maybe_unused]] const char *xb = reinterpret_cast<const char *>(a.x + sub * 4);        // lane bases
    [[maybe_unused]] const char *xaltb = reinterpret_cast<const char *>(xalt + sub * 4);
    [[maybe_unused]] const bool any_alt = ALT && altx.n != 0;
    if constexpr (FAST) {       // opaque: keeps ptxas from re-deriving the lane offset per gather
        asm volatile("" : "+l"(xb));
        asm volatile("" : "+l"(xaltb));
    }
    if constexpr (FAST) {
        static_assert(C::R == 4, "the small-graph walk keeps 4 row ends in registers");
#pragma unroll
        for (int r = 0; r < 4; ++r) ends[r] = __shfl_sync(0xffffffffu, my_end, r, G::LANES);
    }
    [[maybe_unused]] auto prepare_tile = [&](const int2 &c, int t) {
        const int e = chunk_beg + t + sub;
        const int row = (e >= ends[0]) + (e >= ends[1]) + (e >= ends[2]) + (e >= ends[3]);
        meta = (t + sub < n_e) ? row : kSkip;
        const int cc = c.x & LGCN_COL_MASK;
        goff = (unsigned)cc * (unsigned)(D * 4) | (altx.has(cc) ? 0x80000000u : 0u);
    };

    // Batches of U gathers issued ahead of the FMAs that consume them.  (A two-deep A/B
    // register pipeline was measured slower: 128 registers -> 16 warps/SM, and ptxas rotated
    // the loads through temporaries.  The gather microbenchmark profiles/micro/gather_bw.cu
    // shows ~32 KB in flight per SM already saturates HBM for 512-byte random rows.)
    constexpr int U = C::U;
    float4 x[U][G::VEC];
    bool xlive[U];                                        // XF: gather u of the batch read a flagged row
    unsigned touched = 0;                                 // XF: rows of the chunk that summed one
#pragma unroll
    for (int u = 0; u < U; ++u) {
        xlive[u] = true;
#pragma unroll
        for (int v = 0; v < G::VEC; ++v) x[u][v] = zero4;
    }

    bool prefetched = false;
    for (int t = 0; t < max_n; t += G::LANES) {
        if (HINT && MODE != LGCN_SPMM_PLAIN && !prefetched && t + 2 * G::LANES >= max_n) {   // warp-uniform
            prefetched = true;
            if (sub == 0) prefetch_epilogue_rows<D, MODE>(a, r0, nvr);
        }
        int2 cvn = make_int2(0, 0);
        if (t + G::LANES + sub < n_e) cvn = ld_cv<HINT>(cvp + t + G::LANES + sub, pol);   // next tile, one ahead
        const int cnt = min(n_e - t, G::LANES);           // entries of this tile (may be <= 0)
        const int maxcnt = min(G::LANES, max_n - t);      // warp-uniform
        if constexpr (FAST) prepare_tile(cv, t);
        for (int j = 0; j < maxcnt; j += U) {
            // Unconditional loads: slots past the end carry col 0 (a valid, cache-resident row).  A
            // predicated 128-bit load makes ptxas stage through 4 temporaries and serialises the
            // batch (ncu: stalls on the predicated MOVs).
            if (XF) {
                // sparse-input hop (x = g' of the first backward hop): rows flagged all-zero are
                // redirected to the cache-resident zero row; all flag loads of the batch first
                int cjs[U];
                unsigned xfl[U];
                bool live = false;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    cjs[u] = __shfl_sync(0xffffffffu, cv.x, j + u, G::LANES) & LGCN_COL_MASK;
                    xfl[u] = (unsigned)__ldg(a.x_rowflag + cjs[u]);
                    live |= xfl[u] != 0 && j + u < cnt;
                    xlive[u] = xfl[u] != 0;
                }
                // no flagged row among the batch's gathers (the common case while the gradient
                // is still sparse): nothing to add -- rows that never see a live batch are zero
                // filled (or, with y_rowflag, reported as all-zero and not written at all)
                if (!__any_sync(0xffffffffu, live)) continue;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float *src = xfl[u] ? a.x + (size_t)cjs[u] * D + sub * 4 : a.zero_row + sub * 4;
#pragma unroll
                    for (int v = 0; v < G::VEC; ++v) x[u][v] = ld_nc_f4(src + v * G::LANES * 4);
                }
            } else if constexpr (FAST) {
                if (any_alt) {                            // kernel-uniform
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const unsigned o = __shfl_sync(0xffffffffu, goff, j + u, G::LANES);
                        const float4 *src = reinterpret_cast<const float4 *>(
                            ((int)o < 0 ? xaltb : xb) + (o & 0x7fffffffu));
#pragma unroll
                        for (int v = 0; v < G::VEC; ++v) x[u][v] = __ldg(src + v * G::LANES);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const float4 *src = reinterpret_cast<const float4 *>(
                            xb + __shfl_sync(0xffffffffu, goff, j + u, G::LANES));
#pragma unroll
                        for (int v = 0; v < G::VEC; ++v) x[u][v] = __ldg(src + v * G::LANES);
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int cr = __shfl_sync(0xffffffffu, cv.x, j + u, G::LANES);
                    const int cc = cr & LGCN_COL_MASK;
                    const float *src = (altx.has(cc) ? xalt : a.x) + (size_t)cc * D + sub * 4;
                    if (HINT) {
                        const uint64_t gp = pick_policy(gpol, cr);
#pragma unroll
                        for (int v = 0; v < G::VEC; ++v) x[u][v] = ld_nc_f4_hint(src + v * G::LANES * 4, gp);
                    } else {
#pragma unroll
                        for (int v = 0; v < G::VEC; ++v) x[u][v] = ld_nc_f4(src + v * G::LANES * 4);
                    }
                }
            }
            auto flush_to = [&](int row) {                // store the finished row, zero the empty ones
#pragma unroll
                for (int v = 0; v < G::VEC; ++v) {
                    st_f4(stage + cur * D + sub * 4 + v * G::LANES * 4, acc[v]);
                    acc[v] = zero4;
                }
                for (int r = cur + 1; r < row; ++r)
#pragma unroll
                    for (int v = 0; v < G::VEC; ++v)
                        st_f4(stage + r * D + sub * 4 + v * G::LANES * 4, zero4);
                cur = row;
            };
            if constexpr (FAST) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float wj = __int_as_float(__shfl_sync(0xffffffffu, cv.y, j + u, G::LANES));
                    const int row = __shfl_sync(0xffffffffu, meta, j + u, G::LANES);
                    bool take = true;
                    if (row != cur) {                     // group-uniform; rare
                        if (row == kSkip) take = false; else flush_to(row);
                    }
                    if (take) {
                        if (XF && xlive[u]) touched |= 1u << row;
#pragma unroll
                        for (int v = 0; v < G::VEC; ++v) fma4(acc[v], wj, x[u][v]);
                    }
                }
                continue;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float wj = __int_as_float(__shfl_sync(0xffffffffu, cv.y, j + u, G::LANES));
                const int e = chunk_beg + t + j + u;
                // rows of this chunk that end at or before e (one ballot, no per-row pointer chase)
                const unsigned passed = (__ballot_sync(0xffffffffu, my_end <= e) >> gshift) & gbits;
                if (j + u < cnt) {
                    const int row = __popc(passed);
                    if (row != cur) {                     // flush the finished row, zero the empty ones
#pragma unroll
                        for (int v = 0; v < G::VEC; ++v) {
                            st_f4(stage + cur * D + sub * 4 + v * G::LANES * 4, acc[v]);
                            acc[v] = zero4;
                        }
                        for (int r = cur + 1; r < row; ++r)
#pragma unroll
                            for (int v = 0; v < G::VEC; ++v)
                                st_f4(stage + r * D + sub * 4 + v * G::LANES * 4, zero4);
                        cur = row;
                    }
                    if (XF && xlive[u]) touched |= 1u << row;
#pragma unroll
                    for (int v = 0; v < G::VEC; ++v) fma4(acc[v], wj, x[u][v]);
                }
            }
        }
        cv = cvn;
    }
    if (HINT && MODE != LGCN_SPMM_PLAIN && !prefetched && sub == 0) prefetch_epilogue_rows<D, MODE>(a, r0, nvr);
    // rows cur .. nvr-1: the last summed row, then trailing empty rows
    if (cur < nvr) {
#pragma unroll
        for (int v = 0; v < G::VEC; ++v) st_f4(stage + cur * D + sub * 4 + v * G::LANES * 4, acc[v]);
        for (int r = cur + 1; r < nvr; ++r)
#pragma unroll
            for (int v = 0; v < G::VEC; ++v) st_f4(stage + r * D + sub * 4 + v * G::LANES * 4, zero4);
    }
    __syncwarp();
    const unsigned wmask = XF ? sparse_out_mask<C::R>(a, rfw, touched, r0, sub, nvr, my_long) : 0xffffffffu;
    chunk_epilogue<D, MODE, C::R, HINT, ALT>(a, stage, r0, nvr, long_bits, pol, rfw, wmask);
}

// compile-time unrolled loop: f(std::integral_constant<int, I>) for I in [0, N)
template <int N, int I = 0, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<N, I + 1>(f);
    }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_hint(uint32_t dst, const void *src, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;"
                 :: "r"(dst), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// ---- ring kernel: the chunk walk of spmm_chunk_kernel with the gathers in a cp.async ring -------
// One worker per chunk of R rows (R = 8: many small chunks, so that the rowptr -> {col,val} ->
// gather start-up chain of one warp overlaps the streaming of the 20+ others on the SM), rows
// staged in shared memory for the batched epilogue.  Measured at the Amazon shape (d = 128):
// 4.36 ms at 83 % of the DRAM peak (ncu) against 6.0 ms / 60 % for the register-batch kernel.
#ifndef LGCN_RING_WARPS
#define LGCN_RING_WARPS 1
#endif
#ifndef LGCN_RING_S
#define LGCN_RING_S 8
#endif
#ifndef LGCN_RING_R
#define LGCN_RING_R 8
#endif
#ifndef LGCN_RING_S_NARROW
#define LGCN_RING_S_NARROW 8           // ring slots of tables narrower than LGCN_RING_S lanes (d = 16)
#endif
constexpr int kRingWarps = LGCN_RING_WARPS;

template <int D>
struct RingCfg {
    using G = RowGeom<D>;
    static constexpr int R = G::LANES < LGCN_RING_R ? G::LANES : LGCN_RING_R;   // one row end per lane
    // ring slots (power of 2).  A worker of a narrow table has few lanes (4 at d = 16), and with
    // one {col,val} per lane per tile the ring could not be deeper than that: 3 gathers of 64 bytes
    // in flight per worker left the DRAM pipe half empty (ncu r01: 54 % DRAM, 30 warps / SM = the
    // 32-CTA limit).  So the tile is decoupled from the lane count: TL stream entries per tile,
    // EPL = TL / LANES of them per lane, and the ring is S deep whatever the width.
    static constexpr int S = G::LANES < LGCN_RING_S ? LGCN_RING_S_NARROW : LGCN_RING_S;
    static constexpr int TL = G::LANES > S ? G::LANES : S;                       // entries per tile
    static constexpr int EPL = TL / G::LANES;                                    // entries per lane
    static constexpr int SL = G::LANES < S ? G::LANES : S;                       // live-list batch
    static constexpr int WORKERS = kRingWarps * G::GROUPS;
    static constexpr int ROWS_PER_CTA = WORKERS * R;
    static constexpr size_t SMEM = (size_t)WORKERS * (R + S) * D * sizeof(float);
    static_assert((S & (S - 1)) == 0 && S >= 2 && TL % S == 0 && TL % G::LANES == 0, "ring depth");
    static_assert(R % 4 == 0, "chunk rows");
};

// The live-list kernel spends its time in the rowptr -> {col,val} -> flag start-up chain of a chunk,
// not in gathers: more rows per worker amortise that chain (LGCN_LIVE_R rows, one row end per lane).
#ifndef LGCN_LIVE_R
#define LGCN_LIVE_R 8
#endif
template <int D>
struct LiveCfg {
    using G = RowGeom<D>;
    static constexpr int R = G::LANES < LGCN_LIVE_R ? G::LANES : LGCN_LIVE_R;
    static constexpr int S = RingCfg<D>::SL;
    static constexpr int WORKERS = kRingWarps * G::GROUPS;
    static constexpr int ROWS_PER_CTA = WORKERS * R;
    static constexpr size_t SMEM = (size_t)WORKERS * (R + S) * D * sizeof(float);
    static_assert(R % 4 == 0, "chunk rows");
};

template <int D, int MODE, bool HINT, bool ALT, int NLM = 8>
__global__ void __launch_bounds__(kRingWarps * 32)
spmm_ring_kernel(const __grid_constant__ lgcn_spmm_args a) {
    using G = RowGeom<D>;
    using C = RingCfg<D>;
    constexpr int L = G::LANES, S = C::S, TL = C::TL, EPL = C::EPL;
    const uint64_t pol = HINT ? policy_evict_first() : 0ull;
    // per-gather L2 classes only where they were measured to matter (512-byte rows and wider:
    // profiles/r01_hot_columns_sweep.txt shows no effect at d = 16 / 32, where the walk is issue bound)
    constexpr bool GHINT = HINT && D >= 64;
    const GatherPolicy gpol = gather_policy<GHINT>(a.flags, pol);
    extern __shared__ __align__(16) float ring_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int grp = lane / L;
    const int sub = lane % L;
    const int gshift = grp * L;
    const unsigned gbits = (L == 32) ? 0xffffffffu : ((1u << L) - 1u);
    float *stage = ring_smem + (size_t)((warp * G::GROUPS + grp) * (C::R + S)) * D;
    const float *ring = stage + C::R * D + sub * 4;                   // this lane's 16 bytes of slot 0
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);

    const int64_t worker = ((int64_t)blockIdx.x * kRingWarps + warp) * G::GROUPS + grp;
    int64_t chunk = worker;         // optional plan: chunks of similar length share a warp (GROUPS > 1)
    if (G::GROUPS > 1 && a.chunk_order && worker * C::R < a.n_rows) chunk = __ldg(a.chunk_order + worker);
    const int64_t r0 = chunk * C::R;
    const int64_t left = a.n_rows - r0;
    const int nvr = left <= 0 ? 0 : (left < C::R ? (int)left : C::R);

    unsigned rfw[C::R / 4];
#pragma unroll
    for (int i = 0; i < C::R / 4; ++i)
        rfw[i] = ((MODE == LGCN_SPMM_ADD || MODE == LGCN_SPMM_ADAM) && a.addend_rowflag && nvr > 0)
                     ? __ldg(reinterpret_cast<const unsigned *>(a.addend_rowflag + r0) + i) : 0xffffffffu;
    unsigned rb = 0, re = 0;
    if (sub < nvr) {
        rb = __ldg(a.rowptr + r0 + sub);
        re = __ldg(a.rowptr + r0 + sub + 1);
    }
    const bool my_long = (rb >> 31) != 0;
    const int my_beg = (int)(rb & 0x7fffffffu);
    int my_end = (int)(re & 0x7fffffffu);
    int chunk_beg = __shfl_sync(0xffffffffu, my_beg, 0, L);
    int chunk_end = __shfl_sync(0xffffffffu, my_end, nvr > 0 ? nvr - 1 : 0, L);
    if (nvr == 0) chunk_beg = chunk_end = 0;
    const unsigned long_bits = (__ballot_sync(0xffffffffu, my_long) >> gshift) & gbits;
    my_end = sub < nvr ? my_end - chunk_beg : INT_MAX;   // relative to the chunk stream; sentinel

    const int n_e = chunk_end - chunk_beg;
    int max_n = n_e;
#pragma unroll
    for (int off = L; off < 32; off <<= 1)
        max_n = max(max_n, __shfl_xor_sync(0xffffffffu, max_n, off));

    const char *xb = reinterpret_cast<const char *>(a.x + sub * 4);
    const char *xab = reinterpret_cast<const char *>(alt_shifted<D>(a) + sub * 4);
    const AltRange altx = alt_x_range<ALT>(a);
    const int2 *cvp = reinterpret_cast<const int2 *>(a.colval) + chunk_beg;
    const int2 z2 = make_int2(0, 0);
    // {col,val} tiles of TL stream entries: lane `sub` holds entries sub*EPL .. sub*EPL+EPL-1
    struct Tile { int2 e[EPL]; };
    auto load_tile = [&](int t0) {
        Tile T;
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            const int i = t0 + sub * EPL + k;
            T.e[k] = i < n_e ? ld_cv<HINT>(cvp + i, pol) : z2;
        }
        return T;
    };
    Tile cvA = load_tile(0);                                                 // entries [t, t+TL)
    Tile cvB = load_tile(TL);                                                // [t+TL, t+2TL)
    Tile cvC = load_tile(2 * TL);                                            // [t+2TL, t+3TL)
    // Gather of stream entry t+JJ (tile-relative index JJ is a compile-time constant, so the source
    // lane of the shuffle, the tile register and the ring slot are all immediates).
    auto issue = [&](int t, auto jj_c) {
        constexpr int JJ = decltype(jj_c)::value;
        constexpr int JT = JJ % TL;
        const int cr = __shfl_sync(0xffffffffu, JJ < TL ? cvA.e[JT % EPL].x : cvB.e[JT % EPL].x, JT / EPL, L);
        if (t + JJ < n_e) {
            const int cc = cr & LGCN_COL_MASK;
            const char *src = (altx.has(cc) ? xab : xb) + (uint64_t)(uint32_t)cc * (D * 4);
            const uint32_t dst = ring_s + (uint32_t)(JJ % S) * (D * 4);
            if (GHINT) {
                const uint64_t gp = pick_policy(gpol, cr);
#pragma unroll
                for (int v = 0; v < G::VEC; ++v) cp_async16_hint(dst + v * L * 16, src + v * L * 16, gp);
            } else {
#pragma unroll
                for (int v = 0; v < G::VEC; ++v) cp_async16(dst + v * L * 16, src + v * L * 16);
            }
        }
        cp_async_commit();
    };
    static_for<S - 1>([&](auto i) { issue(0, i); });

    float4 acc[G::VEC];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int v = 0; v < G::VEC; ++v) acc[v] = zero4;
    int cur = 0;                                          // row (within the chunk) being summed
    int cur_end = __shfl_sync(0xffffffffu, my_end, 0, L); // first stream entry past that row

    bool prefetched = false;
    for (int t = 0; t < max_n; t += TL) {                 // cvA = entries [t,t+TL), cvB = [t+TL,t+2TL)
        if (HINT && MODE != LGCN_SPMM_PLAIN && !prefetched && t + 2 * TL >= max_n) {         // warp-uniform
            prefetched = true;
            if (sub == 0) prefetch_epilogue_rows<D, MODE>(a, r0, nvr);
        }
        bool done = false;
        static_for<TL>([&](auto j_c) {
            constexpr int J = decltype(j_c)::value;
            if (done) return;
            const int e = t + J;
            if (e >= max_n) { done = true; return; }      // warp-uniform
            issue(t, std::integral_constant<int, J + S - 1>{});
            cp_async_wait<S - 1>();                       // this lane's bytes of entry e have landed
            const float wj = __int_as_float(__shfl_sync(0xffffffffu, cvA.e[J % EPL].y, J / EPL, L));
            const bool live = e < n_e;

            // leave every row that ends at or before e: the finished sum (zeros for the empty rows
            // that follow it) goes to the staging buffer, one row per trip
            while (G::GROUPS == 1 ? (e >= cur_end) : __any_sync(0xffffffffu, live && e >= cur_end)) {
                const bool cross = G::GROUPS == 1 || (live && e >= cur_end);
                if (cross) {
#pragma unroll
                    for (int v = 0; v < G::VEC; ++v) {
                        st_f4(stage + cur * D + sub * 4 + v * L * 4, acc[v]);
                        acc[v] = zero4;
                    }
                    ++cur;
                }
                cur_end = __shfl_sync(0xffffffffu, my_end, cur, L);
            }
            if (live) {
                const float *slot = ring + (J % S) * D;
#pragma unroll
                for (int v = 0; v < G::VEC; ++v) {
                    const float4 xv = *reinterpret_cast<const float4 *>(slot + v * L * 4);
                    fma4(acc[v], wj, xv);
                }
            }
        });
        cvA = cvB;                                        // consume pointer leaves its tile
        cvB = cvC;
        cvC = load_tile(t + 3 * TL);
    }
    cp_async_wait<0>();
    if (HINT && MODE != LGCN_SPMM_PLAIN && !prefetched && sub == 0) prefetch_epilogue_rows<D, MODE>(a, r0, nvr);
    // rows cur .. nvr-1: the last summed row, then trailing empty rows
#pragma unroll 1
    for (; cur < nvr; ++cur) {
#pragma unroll
        for (int v = 0; v < G::VEC; ++v) {
            st_f4(stage + cur * D + sub * 4 + v * L * 4, acc[v]);
            acc[v] = zero4;
        }
    }
    __syncwarp();
    chunk_epilogue<D, MODE, C::R, HINT, ALT, NLM>(a, stage, r0, nvr, long_bits, pol, rfw);
}

// ---- sparse-input hop: live-list kernel --------------------------------------------------------
// Mode ADD with x_rowflag (the first two Horner hops: x = g' has <= 3*batch non-zero rows, its
// first image ~15 % at the Amazon shape).  The dense ring kernel spends ~40 instructions on every
// entry, which makes it ISSUE bound once the gathers are gone (ncu: 60 % issue-active, 2 % DRAM).
// Here a worker (same chunks, staging and epilogue as the ring kernel) resolves the row flags of a
// whole {col,val} tile at once -- one flag byte per lane, fetched a tile ahead -- ballots the LIVE
// entries and walks only those: up to S gathers per batch into the smem slots (cp.async), the row
// of a live entry found with one ballot over the per-lane row ends.  Dead entries cost nothing;
// rows without a live entry are zero filled (or, with y_rowflag, reported and left unwritten).
// The per-row sum is still a sequential fp32 FMA over the live entries in column order, and a
// skipped entry would have added w * 0 = 0 exactly, so results stay bit-equal.
template <int D, bool HINT>
__global__ void __launch_bounds__(kRingWarps * 32)
spmm_live_kernel(const __grid_constant__ lgcn_spmm_args a) {
    using G = RowGeom<D>;
    using C = LiveCfg<D>;
    constexpr int L = G::LANES, S = C::S;
    const uint64_t pol = HINT ? policy_evict_first() : 0ull;
    extern __shared__ __align__(16) float ring_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int grp = lane / L;
    const int sub = lane % L;
    const int gshift = grp * L;
    const unsigned gbits = (L == 32) ? 0xffffffffu : ((1u << L) - 1u);
    float *stage = ring_smem + (size_t)((warp * G::GROUPS + grp) * (C::R + S)) * D;
    const float *ring = stage + C::R * D + sub * 4;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);

    const int64_t worker = ((int64_t)blockIdx.x * kRingWarps + warp) * G::GROUPS + grp;
    int64_t chunk = worker;         // optional plan: chunks of similar length share a warp (GROUPS > 1)
    if (G::GROUPS > 1 && a.chunk_order && worker * C::R < a.n_rows) chunk = __ldg(a.chunk_order + worker);
    const int64_t r0 = chunk * C::R;
    const int64_t left = a.n_rows - r0;
    const int nvr = left <= 0 ? 0 : (left < C::R ? (int)left : C::R);

    unsigned rfw[C::R / 4];
#pragma unroll
    for (int i = 0; i < C::R / 4; ++i)
        rfw[i] = (a.addend_rowflag && nvr > 0)
                     ? __ldg(reinterpret_cast<const unsigned *>(a.addend_rowflag + r0) + i) : 0xffffffffu;
    unsigned rb = 0, re = 0;
    if (sub < nvr) {
        rb = __ldg(a.rowptr + r0 + sub);
        re = __ldg(a.rowptr + r0 + sub + 1);
    }
    const bool my_long = (rb >> 31) != 0;
    const int my_beg = (int)(rb & 0x7fffffffu);
    int my_end = (int)(re & 0x7fffffffu);
    int chunk_beg = __shfl_sync(0xffffffffu, my_beg, 0, L);
    int chunk_end = __shfl_sync(0xffffffffu, my_end, nvr > 0 ? nvr - 1 : 0, L);
    if (nvr == 0) chunk_beg = chunk_end = 0;
    const unsigned long_bits = (__ballot_sync(0xffffffffu, my_long) >> gshift) & gbits;
    my_end = sub < nvr ? my_end - chunk_beg : INT_MAX;   // relative to the chunk stream; sentinel

    const int n_e = chunk_end - chunk_beg;
    int max_n = n_e;
#pragma unroll
    for (int off = L; off < 32; off <<= 1)
        max_n = max(max_n, __shfl_xor_sync(0xffffffffu, max_n, off));

    const char *xb = reinterpret_cast<const char *>(a.x + sub * 4);
    const int2 *cvp = reinterpret_cast<const int2 *>(a.colval) + chunk_beg;
    const int2 z2 = make_int2(0, 0);
    // tile pipeline: cvA = entries [t,t+L) with its flags resolved, cvB = [t+L,t+2L) with the flag
    // load in flight, cvC = [t+2L,t+3L) in flight
    int2 cvA = sub < n_e ? ld_cv<HINT>(cvp + sub, pol) : z2;
    int2 cvB = L + sub < n_e ? ld_cv<HINT>(cvp + L + sub, pol) : z2;
    int2 cvC = 2 * L + sub < n_e ? ld_cv<HINT>(cvp + 2 * L + sub, pol) : z2;
    unsigned flA = __ldg(a.x_rowflag + (cvA.x & LGCN_COL_MASK));
    unsigned flB = __ldg(a.x_rowflag + (cvB.x & LGCN_COL_MASK));

    float4 acc[G::VEC];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int v = 0; v < G::VEC; ++v) acc[v] = zero4;
    int cur = 0;                                          // row (within the chunk) being summed
    unsigned touched = 0;                                 // rows that summed a live entry

    for (int t = 0; t < max_n; t += L) {
        const bool mine = t + sub < n_e && flA != 0;
        unsigned lm = (__ballot_sync(0xffffffffu, mine) >> gshift) & gbits;    // live entries of the tile
        while (__any_sync(0xffffffffu, lm != 0)) {
            // up to S live entries: gathers first ...
            unsigned m = lm;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const bool on = m != 0;
                const int j = on ? __ffs(m) - 1 : 0;
                m &= m - 1;
                const int cj = __shfl_sync(0xffffffffu, cvA.x, j, L) & LGCN_COL_MASK;
                if (on) {
                    const char *src = xb + (uint64_t)(uint32_t)cj * (D * 4);
                    const uint32_t dst = ring_s + (uint32_t)s * (D * 4);
#pragma unroll
                    for (int v = 0; v < G::VEC; ++v) cp_async16(dst + v * L * 16, src + v * L * 16);
                }
            }
            cp_async_commit();
            cp_async_wait<0>();
            // ... then their FMAs, in stream (= column) order
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const bool on = lm != 0;
                const int j = on ? __ffs(lm) - 1 : 0;
                lm &= lm - 1;
                const float wj = __int_as_float(__shfl_sync(0xffffffffu, cvA.y, j, L));
                const int e = t + j;
                const unsigned passed = (__ballot_sync(0xffffffffu, my_end <= e) >> gshift) & gbits;
                if (on) {
                    const int row = __popc(passed);
                    if (row != cur) {                     // flush the finished row, zero the skipped ones
#pragma unroll
                        for (int v = 0; v < G::VEC; ++v) {
                            st_f4(stage + cur * D + sub * 4 + v * L * 4, acc[v]);
                            acc[v] = zero4;
                        }
                        for (int r = cur + 1; r < row; ++r)
#pragma unroll
                            for (int v = 0; v < G::VEC; ++v)
                                st_f4(stage + r * D + sub * 4 + v * L * 4, zero4);
                        cur = row;
                    }
                    touched |= 1u << row;
                    const float *slot = ring + s * D;
#pragma unroll
                    for (int v = 0; v < G::VEC; ++v) {
                        const float4 xv = *reinterpret_cast<const float4 *>(slot + v * L * 4);
                        fma4(acc[v], wj, xv);
                    }
                }
            }
        }
        cvA = cvB;
        flA = flB;
        cvB = cvC;
        flB = __ldg(a.x_rowflag + (cvB.x & LGCN_COL_MASK));
        cvC = t + 3 * L + sub < n_e ? ld_cv<HINT>(cvp + t + 3 * L + sub, pol) : z2;
    }
    if (cur < nvr) {
#pragma unroll
        for (int v = 0; v < G::VEC; ++v) st_f4(stage + cur * D + sub * 4 + v * L * 4, acc[v]);
        for (int r = cur + 1; r < nvr; ++r)
#pragma unroll
            for (int v = 0; v < G::VEC; ++v) st_f4(stage + r * D + sub * 4 + v * L * 4, zero4);
    }
    __syncwarp();
    const unsigned wmask = sparse_out_mask<C::R>(a, rfw, touched, r0, sub, nvr, my_long);
    chunk_epilogue<D, LGCN_SPMM_ADD, C::R, HINT, false>(a, stage, r0, nvr, long_bits, pol, rfw, wmask);
}

// ---- long rows: one worker per segment, partial sums to seg_ws -----------------------------
// FMODE >= 0 (small graphs, inside the chunk kernel's launch, lgcn_spmm_args.long_done given): the
// worker that delivers the LAST partial of a long row also combines the row and runs its epilogue
// in mode FMODE -- fence, count, fence: the classic last-arriver hand-over, no waiting anywhere.
template <int D, bool ALT, int FMODE>
__device__ __forceinline__ void long_seg_body(const lgcn_spmm_args &a, int64_t block) {
    using G = RowGeom<D>;
    const int lane = threadIdx.x & 31;
    const int grp = lane / G::LANES;
    const int sub = lane % G::LANES;
    const int64_t warp = block * kWarps + (threadIdx.x >> 5);
    const int64_t seg = warp * G::GROUPS + grp;
    int beg = 0, deg = 0;
    [[maybe_unused]] int long_i = 0;                      // the long row this segment belongs to
    if (seg < a.n_seg) {
        int lo = 0, hi = a.n_long;                        // last i with long_seg_ptr[i] <= seg
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(a.long_seg_ptr + mid) <= seg) lo = mid; else hi = mid;
        }
        long_i = lo;
        const int rbeg = __ldg(a.long_rowptr + lo), rend = __ldg(a.long_rowptr + lo + 1);
        beg = rbeg + (int)(seg - __ldg(a.long_seg_ptr + lo)) * a.seg_len;
        deg = min(a.seg_len, rend - beg);
    }
    int maxdeg = deg;
#pragma unroll
    for (int off = G::LANES; off < 32; off <<= 1)
        maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, off));
    float4 acc[G::VEC];
#pragma unroll
    for (int v = 0; v < G::VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int2 *cvp = reinterpret_cast<const int2 *>(a.long_colval) + beg;
    const AltRange altx = alt_x_range<ALT>(a);
    const float *xalt = alt_shifted<D>(a);
    if (a.x_rowflag) {
        // flagged input (the sparse backward hops): one flag byte per lane for the whole tile, a
        // ballot of the live entries, and only those are gathered and summed -- in stream order,
        // so the partial sum is bit-equal to the dense walk (a dead entry adds w * 0)
        const int gshift = grp * G::LANES;
        const unsigned gbits = (G::LANES == 32) ? 0xffffffffu : ((1u << G::LANES) - 1u);
        for (int base = 0; base < maxdeg; base += G::LANES) {
            int2 cv = make_int2(0, 0);
            unsigned fl = 0;
            if (base + sub < deg) {
                cv = __ldg(cvp + base + sub);
                fl = __ldg(a.x_rowflag + cv.x);
            }
            unsigned lm = (__ballot_sync(0xffffffffu, fl != 0) >> gshift) & gbits;
            while (__any_sync(0xffffffffu, lm != 0)) {
                float4 x[kUnroll][G::VEC];
                unsigned m = lm;
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const bool on = m != 0;
                    const int j = on ? __ffs(m) - 1 : 0;
                    m &= m - 1;
                    const int cj = __shfl_sync(0xffffffffu, cv.x, j, G::LANES);
                    const float *src = on ? a.x + (size_t)cj * D + sub * 4 : a.zero_row + sub * 4;
#pragma unroll
                    for (int v = 0; v < G::VEC; ++v) x[u][v] = ld_nc_f4(src + v * G::LANES * 4);
                }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const bool on = lm != 0;
                    const int j = on ? __ffs(lm) - 1 : 0;
                    lm &= lm - 1;
                    const float wj = __int_as_float(__shfl_sync(0xffffffffu, cv.y, j, G::LANES));
                    if (on) {
#pragma unroll
                        for (int v = 0; v < G::VEC; ++v) fma4(acc[v], wj, x[u][v]);
                    }
                }
            }
        }
    } else
    for (int base = 0; base < maxdeg; base += G::LANES) {
        int2 cv = make_int2(0, 0);
        if (base + sub < deg) cv = __ldg(cvp + base + sub);
        const int cnt = min(deg - base, G::LANES);
        const int maxcnt = min(G::LANES, maxdeg - base);
        for (int j = 0; j < maxcnt; j += kUnroll) {
            float4 x[kUnroll][G::VEC];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
#pragma unroll
                for (int v = 0; v < G::VEC; ++v) x[u][v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int cj = __shfl_sync(0xffffffffu, cv.x, j + u, G::LANES);
                const float *src = (altx.has(cj) ? xalt : a.x) + (size_t)cj * D + sub * 4;
#pragma unroll
                for (int v = 0; v < G::VEC; ++v) x[u][v] = ld_nc_f4(src + v * G::LANES * 4);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const float wj = __int_as_float(__shfl_sync(0xffffffffu, cv.y, j + u, G::LANES));
                if (j + u < cnt) {
#pragma unroll
                    for (int v = 0; v < G::VEC; ++v) fma4(acc[v], wj, x[u][v]);
                }
            }
        }
    }
    if (seg < a.n_seg) {
#pragma unroll
        for (int v = 0; v < G::VEC; ++v)
            st_f4(a.seg_ws + (size_t)seg * D + sub * 4 + v * G::LANES * 4, acc[v]);
    }
    if constexpr (FMODE >= 0) {
        if (a.long_done) {                                // kernel-uniform
            if (seg < a.n_seg) __threadfence();           // my part of the partial is visible device-wide
            __syncwarp();
            int old = -2;
            if (seg < a.n_seg && sub == 0) old = atomicAdd(a.long_done + long_i, 1);
            old = __shfl_sync(0xffffffffu, old, 0, G::LANES);
            if (seg < a.n_seg &&
                old + 1 == __ldg(a.long_seg_ptr + long_i + 1) - __ldg(a.long_seg_ptr + long_i)) {
                __threadfence();                          // the other workers' partials, after their counts
                combine_long_row<D, FMODE, ALT, 8 / RowGeom<D>::VEC>(a, long_i, sub);
                if (sub == 0) a.long_done[long_i] = 0;    // re-armed for the next call
            }
        }
    }
}

template <int D, bool ALT>
__global__ void __launch_bounds__(kThreads) spmm_long_seg_kernel(const __grid_constant__ lgcn_spmm_args a) {
    long_seg_body<D, ALT>(a, (int64_t)blockIdx.x);
}

// ---- long rows: combine the segment partials in order, then the epilogue -------------------
// (partials come from other SMs: L2 loads, never a stale L1 line)
template <int D, int MODE, bool ALT, int CBMAX>      // CBMAX: partial loads in flight (8 inside the 64-register chunk kernel)
__device__ __forceinline__ void combine_long_row(const lgcn_spmm_args &a, int i, int sub) {
    using G = RowGeom<D>;
    const int s0 = __ldg(a.long_seg_ptr + i), s1 = __ldg(a.long_seg_ptr + i + 1);
    float4 acc[G::VEC];
#pragma unroll
    for (int v = 0; v < G::VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    // partials are summed in segment order; loads are issued CB segments ahead of the adds and are
    // unconditional (clamped index): the hottest row's chain of dependent batches is the whole
    // duration of a separate combine launch on small graphs (Gowalla shape: ~300 segments)
    constexpr int CB = (G::VEC > 1 ? 8 : 16) < CBMAX ? (G::VEC > 1 ? 8 : 16) : CBMAX;
    for (int s = s0; s < s1; s += CB) {
        float4 t[CB][G::VEC];
#pragma unroll
        for (int k = 0; k < CB; ++k) {
            const int sk = min(s + k, s1 - 1);
#pragma unroll
            for (int v = 0; v < G::VEC; ++v)
                t[k][v] = __ldcg(reinterpret_cast<const float4 *>(a.seg_ws + (size_t)sk * D + sub * 4 +
                                                                  v * G::LANES * 4));
        }
#pragma unroll
        for (int k = 0; k < CB; ++k)
            if (s + k < s1) {
#pragma unroll
                for (int v = 0; v < G::VEC; ++v) {
                    if (s + k == s0) acc[v] = t[k][v]; else add4(acc[v], t[k][v]);
                }
            }
    }
    epilogue_row<D, MODE, ALT>(a, __ldg(a.long_row_ids + i), acc);
}

template <int D, int MODE, bool ALT>
__global__ void __launch_bounds__(kThreads) spmm_long_combine_kernel(const __grid_constant__ lgcn_spmm_args a) {
    using G = RowGeom<D>;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int64_t i = warp * G::GROUPS + lane / G::LANES;
    if (i >= a.n_long) return;
    combine_long_row<D, MODE, ALT, 16>(a, (int)i, lane % G::LANES);
}

template <int D, int MODE, int RSEL, bool HINT, bool XF, bool ALT>
static int launch_chunks(const lgcn_spmm_args &a, cudaStream_t st) {
    using C = ChunkCfg<D, RSEL>;
    LGCN_OPT_IN_SMEM((spmm_chunk_kernel<D, MODE, RSEL, HINT, XF, ALT>), C::SMEM);
    int64_t gb = (a.n_rows + C::ROWS_PER_CTA - 1) / C::ROWS_PER_CTA;
    if (RSEL == 1 && a.n_long > 0)          // + the long-row segment workers (see the kernel)
        gb += (a.n_seg + kWarps * RowGeom<D>::GROUPS - 1) / (kWarps * RowGeom<D>::GROUPS);
    if (gb > 0x7fffffffLL) return LGCN_E_TOO_LARGE;
    spmm_chunk_kernel<D, MODE, RSEL, HINT, XF, ALT><<<(unsigned)gb, kThreads, C::SMEM, st>>>(a);
    LGCN_LAUNCH_CHECK();
    return 0;
}

template <int D, bool HINT>
static int launch_live(const lgcn_spmm_args &a, cudaStream_t st) {
    using C = LiveCfg<D>;
    LGCN_OPT_IN_SMEM((spmm_live_kernel<D, HINT>), C::SMEM);
    const int64_t gb = (a.n_rows + C::ROWS_PER_CTA - 1) / C::ROWS_PER_CTA;
    if (gb > 0x7fffffffLL) return LGCN_E_TOO_LARGE;
    spmm_live_kernel<D, HINT><<<(unsigned)gb, kRingWarps * 32, C::SMEM, st>>>(a);
    LGCN_LAUNCH_CHECK();
    return 0;
}

template <int D, int MODE, bool HINT, bool ALT>
static int launch_ring(const lgcn_spmm_args &a, cudaStream_t st) {
    using C = RingCfg<D>;
    const int64_t gb = (a.n_rows + C::ROWS_PER_CTA - 1) / C::ROWS_PER_CTA;
    if (gb > 0x7fffffffLL) return LGCN_E_TOO_LARGE;
    if (MODE == LGCN_SPMM_MEAN && a.n_layers <= 4) {        // K <= 4: the lean MEAN epilogue
        LGCN_OPT_IN_SMEM((spmm_ring_kernel<D, MODE, HINT, ALT, 4>), C::SMEM);
        spmm_ring_kernel<D, MODE, HINT, ALT, 4><<<(unsigned)gb, kRingWarps * 32, C::SMEM, st>>>(a);
    } else {
        LGCN_OPT_IN_SMEM((spmm_ring_kernel<D, MODE, HINT, ALT, 8>), C::SMEM);
        spmm_ring_kernel<D, MODE, HINT, ALT, 8><<<(unsigned)gb, kRingWarps * 32, C::SMEM, st>>>(a);
    }
    LGCN_LAUNCH_CHECK();
    return 0;
}

// small graphs: 4-row chunks so that the chip is filled (>= ~2 waves of workers), and the long-row
// segment workers ride in the chunk launch
template <int D>
static bool small_graph(int64_t n_rows, int32_t flags) {
    const int64_t big_workers = n_rows / ChunkCfg<D, 0>::R;
    return big_workers < (int64_t)kNumSMs * 32 * RowGeom<D>::GROUPS && !(flags & LGCN_SPMM_F_BIG_PATH);
}

// Rows per chunk of the main kernel lgcn_spmm will run, when that kernel follows chunk_order (0 = it
// does not: one worker per warp, or the register-batch chunk kernel of the large-graph path).
template <int D>
static int order_chunk_rows(int64_t n_rows, int32_t flags) {
    if (small_graph<D>(n_rows, flags)) return 4;
    if (flags & LGCN_SPMM_F_NO_RING) return 0;
    return RowGeom<D>::GROUPS > 1 ? RingCfg<D>::R : 0;      // ring and live kernels: same chunks
}

// Large graphs: the cp.async ring kernel for every epilogue; the flagged (sparse-input) hops run
// the live-list kernel.  (Round 1 kept the register-batch chunk kernel for ADAM; re-measured in
// round 2 the ring kernel wins at every width: d=16 2.95 -> 2.44 ms, d=32 4.34 -> 3.45, d=64
// 6.77 -> 5.74, d=128 12.18 -> 11.49 -- profiles/r02_adam_ring.txt.  LGCN_SPMM_F_NO_RING still
// selects the chunk kernel; LGCN_SPMM_F_FORCE_RING is kept for ABI compatibility and is a no-op.)
static bool ring_path(bool small, int mode, int32_t flags) {
    (void)mode;
    return !small && !(flags & LGCN_SPMM_F_NO_RING);
}

// ALT: the call reads rows from x_alt (LGCN_SPMM_F_ALT_X / ALT_LAYER0; lgcn_spmm() admits the flags
// for PLAIN, dense ADD and MEAN only) or, in ADAM mode, carries g_skip -- its own instantiations, so
// that every other call pays nothing.
template <int D, int MODE, bool ALT>
static int launch_mode_alt(const lgcn_spmm_args &a, cudaStream_t st) {
    using G = RowGeom<D>;
    constexpr int groups_per_block = kWarps * G::GROUPS;
    const bool small = small_graph<D>(a.n_rows, a.flags);
    if (a.n_long > 0 && !(small && a.n_rows > 0)) {     // small graphs: inside the chunk launch
        const unsigned gs = (unsigned)((a.n_seg + groups_per_block - 1) / groups_per_block);
        spmm_long_seg_kernel<D, ALT><<<gs, kThreads, 0, st>>>(a);
        LGCN_LAUNCH_CHECK();
    }
    if (a.n_rows > 0) {
        const bool hint = (a.flags & LGCN_SPMM_F_STREAM_HINTS) != 0;
        int rc;
        const bool xf = MODE == LGCN_SPMM_ADD && a.x_rowflag != nullptr;
        const bool ring = ring_path(small, MODE, a.flags);
        if (ring && xf) {
            rc = hint ? launch_live<D, true>(a, st) : launch_live<D, false>(a, st);
        } else if (ring) {
            rc = hint ? launch_ring<D, MODE, true, ALT>(a, st) : launch_ring<D, MODE, false, ALT>(a, st);
        } else if (xf) {        // sparse-input hop (first Horner hop): flagged gathers
            if (small) rc = launch_chunks<D, LGCN_SPMM_ADD, 1, false, true, false>(a, st);
            else if (hint) rc = launch_chunks<D, LGCN_SPMM_ADD, 0, true, true, false>(a, st);
            else rc = launch_chunks<D, LGCN_SPMM_ADD, 0, false, true, false>(a, st);
        } else if (small) rc = launch_chunks<D, MODE, 1, false, false, ALT>(a, st);
        else if (hint) rc = launch_chunks<D, MODE, 0, true, false, ALT>(a, st);
        else rc = launch_chunks<D, MODE, 0, false, false, ALT>(a, st);
        if (rc) return rc;
    }
    if (a.n_long > 0 && !(small && a.n_rows > 0 && a.long_done)) {   // small + long_done: combined in place
        const unsigned gc = (unsigned)((a.n_long + groups_per_block - 1) / groups_per_block);
        spmm_long_combine_kernel<D, MODE, ALT><<<gc, kThreads, 0, st>>>(a);
        LGCN_LAUNCH_CHECK();
    }
    return 0;
}

template <int D, int MODE>
static int launch_mode(const lgcn_spmm_args &a, cudaStream_t st) {
    if constexpr (MODE != LGCN_SPMM_ADAM) {
        if (a.flags & (LGCN_SPMM_F_ALT_X | LGCN_SPMM_F_ALT_LAYER0)) return launch_mode_alt<D, MODE, true>(a, st);
    } else {
        if (a.g_skip) return launch_mode_alt<D, MODE, true>(a, st);
    }
    return launch_mode_alt<D, MODE, false>(a, st);
}

template <int D>
static int launch_dim(const lgcn_spmm_args &a, cudaStream_t st) {
    switch (a.mode) {
        case LGCN_SPMM_PLAIN: return launch_mode<D, LGCN_SPMM_PLAIN>(a, st);
        case LGCN_SPMM_ADD:   return launch_mode<D, LGCN_SPMM_ADD>(a, st);
        case LGCN_SPMM_MEAN:  return launch_mode<D, LGCN_SPMM_MEAN>(a, st);
        case LGCN_SPMM_ADAM:  return launch_mode<D, LGCN_SPMM_ADAM>(a, st);
        default: return LGCN_E_BAD_ARG;
    }
}


}  // namespace lgcn

// One translation unit per width instantiates launch_dim<D> behind this plain function.
#define LGCN_SPMM_DEFINE_WIDTH(D)                                                              \
    extern "C" __attribute__((visibility("hidden"))) int lgcn_spmm_launch_d##D(                \
        const lgcn_spmm_args *a, cudaStream_t st) {                                            \
        return lgcn::launch_dim<D>(*a, st);                                                    \
    }
