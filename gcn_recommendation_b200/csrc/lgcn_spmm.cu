// lgcn_spmm.cu -- normalised-adjacency CSR SpMM with fused epilogues (sm_100a).
//
// Replaces torch.sparse.mm (reference models/lightgcn.py:45, models/lightgcn_fusion.py:56),
// its autograd backward (A is symmetric, so the same kernel serves), the layer mean
// (reference models/lightgcn.py:54) and, in ADAM mode, optimizer.step() (reference
// main.py:526).
//
// Mapping: a table row of d floats is owned by a sub-warp group of d/4 lanes (float4 per
// lane); every lane owns fixed feature columns, so one row is a SEQUENTIAL fp32 FMA chain
// in ascending column order -- bit-equal to the CPU reference -- while the loads of the
// gathered rows are issued UNROLL deep ahead of the FMAs.  HBM-bound: no tensor cores.
#include "lgcn_common.cuh"

namespace lgcn {

struct SpmmParams {
    lgcn_spmm_args a;
};

constexpr int kUnroll = 8;
constexpr int kThreads = 256;

// Accumulate entries [beg, beg+deg) of one CSR row into acc for the calling group.
// All 32 lanes of the warp must call this together (loop bounds are made warp-uniform).
template <int D>
__device__ __forceinline__ void accumulate_range(const int32_t *__restrict__ col,
                                                 const float *__restrict__ val,
                                                 const float *__restrict__ X, int beg, int deg,
                                                 float4 (&acc)[RowGeom<D>::VEC]) {
    using G = RowGeom<D>;
    const int lane = threadIdx.x & 31;
    const int sub = lane % G::LANES;
    int maxdeg = deg;
#pragma unroll
    for (int off = G::LANES; off < 32; off <<= 1)
        maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, off));

    for (int base = 0; base < maxdeg; base += G::LANES) {
        int c = 0;
        float w = 0.0f;
        if (base + sub < deg) {
            c = ld_stream_i32(col + beg + base + sub);
            w = ld_stream_f32(val + beg + base + sub);
        }
        const int cnt = min(G::LANES, deg - base);         // may be <= 0 for a finished group
        const int maxcnt = min(G::LANES, maxdeg - base);   // warp-uniform
        for (int j = 0; j < maxcnt; j += kUnroll) {
            float4 x[kUnroll][G::VEC];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int cj = __shfl_sync(0xffffffffu, c, j + u, G::LANES);
                if (j + u < cnt) {
                    const float *src = X + (size_t)cj * D + sub * 4;
#pragma unroll
                    for (int v = 0; v < G::VEC; ++v) x[u][v] = ld_nc_f4(src + v * G::LANES * 4);
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const float wj = __shfl_sync(0xffffffffu, w, j + u, G::LANES);
                if (j + u < cnt) {
#pragma unroll
                    for (int v = 0; v < G::VEC; ++v) fma4(acc[v], wj, x[u][v]);
                }
            }
        }
    }
}

// Apply the epilogue for local row `row` (group-cooperative; each lane owns 4*VEC columns).
template <int D, int MODE>
__device__ __forceinline__ void epilogue(const lgcn_spmm_args &a, int64_t row,
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
            // sequential sum E_0 + E_1 + ... + E_{n-1} + (A x), then a true division
            float4 s = ld_stream_f4(a.layers[0] + off);
            for (int l = 1; l < a.n_layers; ++l) {
                const float4 t = ld_stream_f4(a.layers[l] + off);
                add4(s, t);
            }
            add4(s, acc[v]);
            const float div = (float)(a.n_layers + 1);
            s.x = __fdiv_rn(s.x, div);
            s.y = __fdiv_rn(s.y, div);
            s.z = __fdiv_rn(s.z, div);
            s.w = __fdiv_rn(s.w, div);
            st_f4(a.y + off, s);
        } else {  // LGCN_SPMM_ADAM
            float4 g = acc[v];
            if (a.addend) {
                const float4 t = ld_stream_f4(a.addend + off);
                add4(g, t);
            }
            if (a.addend2) {
                const float4 t = ld_stream_f4(a.addend2 + off);
                add4(g, t);
            }
            float4 p = *reinterpret_cast<const float4 *>(a.p + off);
            float4 m = *reinterpret_cast<const float4 *>(a.m + off);
            float4 vv = *reinterpret_cast<const float4 *>(a.v + off);
            const float ss = __ldg(a.adam_scalars), bs = __ldg(a.adam_scalars + 1);
            adam4(p, m, vv, g, ss, bs, a.beta1, a.beta2, a.eps);
            st_f4(a.p + off, p);
            st_f4(a.m + off, m);
            st_f4(a.v + off, vv);
            if (a.g_out) st_f4(a.g_out + off, g);
        }
    }
}

// ---- main kernel: one group per row ----------------------------------------------------
template <int D, int MODE>
__global__ void __launch_bounds__(kThreads) spmm_rows_kernel(const SpmmParams p) {
    using G = RowGeom<D>;
    const lgcn_spmm_args &a = p.a;
    const int lane = threadIdx.x & 31;
    const int grp = lane / G::LANES;
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t row = warp * G::GROUPS + grp;
    int beg = 0, deg = 0;
    bool owner = false;
    if (row < a.n_rows) {
        beg = __ldg(a.rowptr + row);
        deg = __ldg(a.rowptr + row + 1) - beg;
        owner = true;
        if (a.long_row_threshold > 0 && deg > a.long_row_threshold) {
            deg = 0;        // handled by the segment kernels
            owner = false;
        }
    }
    float4 acc[G::VEC];
#pragma unroll
    for (int v = 0; v < G::VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    accumulate_range<D>(a.col, a.val, a.x, beg, deg, acc);
    if (owner) epilogue<D, MODE>(a, row, acc);
}

// ---- long rows: one group per segment, partial sums to seg_ws ---------------------------
template <int D>
__global__ void __launch_bounds__(kThreads) spmm_long_seg_kernel(const SpmmParams p) {
    using G = RowGeom<D>;
    const lgcn_spmm_args &a = p.a;
    const int lane = threadIdx.x & 31;
    const int grp = lane / G::LANES;
    const int sub = lane % G::LANES;
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t seg = warp * G::GROUPS + grp;
    int beg = 0, deg = 0;
    if (seg < a.n_seg) {
        // long row that owns this segment: last i with long_seg_ptr[i] <= seg
        int lo = 0, hi = a.n_long;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(a.long_seg_ptr + mid) <= seg) lo = mid; else hi = mid;
        }
        const int row = __ldg(a.long_row_ids + lo);
        const int rbeg = __ldg(a.rowptr + row), rend = __ldg(a.rowptr + row + 1);
        beg = rbeg + (int)(seg - __ldg(a.long_seg_ptr + lo)) * a.seg_len;
        deg = min(a.seg_len, rend - beg);
    }
    float4 acc[G::VEC];
#pragma unroll
    for (int v = 0; v < G::VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    accumulate_range<D>(a.col, a.val, a.x, beg, deg, acc);
    if (seg < a.n_seg) {
#pragma unroll
        for (int v = 0; v < G::VEC; ++v)
            st_f4(a.seg_ws + (size_t)seg * D + sub * 4 + v * G::LANES * 4, acc[v]);
    }
}

// ---- long rows: combine the segment partials in order, then the epilogue ----------------
template <int D, int MODE>
__global__ void __launch_bounds__(kThreads) spmm_long_combine_kernel(const SpmmParams p) {
    using G = RowGeom<D>;
    const lgcn_spmm_args &a = p.a;
    const int lane = threadIdx.x & 31;
    const int grp = lane / G::LANES;
    const int sub = lane % G::LANES;
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t i = warp * G::GROUPS + grp;
    if (i >= a.n_long) return;
    const int s0 = __ldg(a.long_seg_ptr + i), s1 = __ldg(a.long_seg_ptr + i + 1);
    float4 acc[G::VEC];
#pragma unroll
    for (int v = 0; v < G::VEC; ++v)
        acc[v] = *reinterpret_cast<const float4 *>(a.seg_ws + (size_t)s0 * D + sub * 4 +
                                                   v * G::LANES * 4);
    for (int s = s0 + 1; s < s1; ++s) {
#pragma unroll
        for (int v = 0; v < G::VEC; ++v) {
            const float4 t = *reinterpret_cast<const float4 *>(a.seg_ws + (size_t)s * D + sub * 4 +
                                                               v * G::LANES * 4);
            add4(acc[v], t);
        }
    }
    epilogue<D, MODE>(a, __ldg(a.long_row_ids + i), acc);
}

template <int D, int MODE>
static int launch_mode(const SpmmParams &p, cudaStream_t st) {
    using G = RowGeom<D>;
    constexpr int rows_per_block = (kThreads / 32) * G::GROUPS;
    const lgcn_spmm_args &a = p.a;
    if (a.long_row_threshold > 0 && a.n_long > 0) {
        const unsigned gs = (unsigned)((a.n_seg + rows_per_block - 1) / rows_per_block);
        spmm_long_seg_kernel<D><<<gs, kThreads, 0, st>>>(p);
        LGCN_LAUNCH_CHECK();
    }
    if (a.n_rows > 0) {
        const int64_t gb = (a.n_rows + rows_per_block - 1) / rows_per_block;
        if (gb > 0x7fffffffLL) return LGCN_E_TOO_LARGE;
        spmm_rows_kernel<D, MODE><<<(unsigned)gb, kThreads, 0, st>>>(p);
        LGCN_LAUNCH_CHECK();
    }
    if (a.long_row_threshold > 0 && a.n_long > 0) {
        const unsigned gc = (unsigned)((a.n_long + rows_per_block - 1) / rows_per_block);
        spmm_long_combine_kernel<D, MODE><<<gc, kThreads, 0, st>>>(p);
        LGCN_LAUNCH_CHECK();
    }
    return 0;
}

template <int D>
static int launch_dim(const SpmmParams &p, cudaStream_t st) {
    switch (p.a.mode) {
        case LGCN_SPMM_PLAIN: return launch_mode<D, LGCN_SPMM_PLAIN>(p, st);
        case LGCN_SPMM_ADD:   return launch_mode<D, LGCN_SPMM_ADD>(p, st);
        case LGCN_SPMM_MEAN:  return launch_mode<D, LGCN_SPMM_MEAN>(p, st);
        case LGCN_SPMM_ADAM:  return launch_mode<D, LGCN_SPMM_ADAM>(p, st);
        default: return LGCN_E_BAD_ARG;
    }
}

}  // namespace lgcn

extern "C" int lgcn_spmm(const lgcn_spmm_args *args, lgcn_stream_t stream) {
    using namespace lgcn;
    if (!args) return LGCN_E_BAD_ARG;
    const lgcn_spmm_args &a = *args;
    if (!dim_supported(a.d)) return LGCN_E_BAD_DIM;
    if (a.n_rows < 0 || !a.rowptr || !a.x) return LGCN_E_BAD_ARG;
    if (a.n_rows > 0x7fffffffLL) return LGCN_E_TOO_LARGE;
    if (a.n_rows > 0 && (!a.col || !a.val)) return LGCN_E_BAD_ARG;
    switch (a.mode) {
        case LGCN_SPMM_PLAIN: if (!a.y) return LGCN_E_BAD_ARG; break;
        case LGCN_SPMM_ADD:   if (!a.y || !a.addend) return LGCN_E_BAD_ARG; break;
        case LGCN_SPMM_MEAN:
            if (!a.y || a.n_layers < 1 || a.n_layers > 8) return LGCN_E_BAD_ARG;
            for (int l = 0; l < a.n_layers; ++l) if (!a.layers[l]) return LGCN_E_BAD_ARG;
            break;
        case LGCN_SPMM_ADAM:
            if (!a.p || !a.m || !a.v || !a.adam_scalars) return LGCN_E_BAD_ARG;
            break;
        default: return LGCN_E_BAD_ARG;
    }
    SpmmParams p;
    p.a = a;
    if (a.long_row_threshold > 0 && a.n_long > 0) {
        if (!a.long_row_ids || !a.long_seg_ptr || !a.seg_ws || a.seg_len <= 0 || a.n_seg <= 0)
            return LGCN_E_BAD_ARG;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (a.d) {
        case 16:  return launch_dim<16>(p, st);
        case 32:  return launch_dim<32>(p, st);
        case 64:  return launch_dim<64>(p, st);
        case 128: return launch_dim<128>(p, st);
        case 256: return launch_dim<256>(p, st);
    }
    return LGCN_E_BAD_DIM;
}

extern "C" size_t lgcn_sizeof_spmm_args(void) { return sizeof(lgcn_spmm_args); }
