// lgcn_fusion.cu -- LightGCN_Fusion item block: side-embedding projection and merge (sm_100a).
//
// Replaces reference models/lightgcn_fusion.py:45-49:
//     H = leaky_relu( cat([E_id, C], 1) @ W.T + b )        W: [d, d+c]
// and its autograd backward.  The [n_items, d+c] concatenation is never materialised: the
// loaders read E_id for k < d and the content matrix C for k >= d.
//
// v1: fp32 SIMT tiles (exact fp32 FMA accumulation).  The tcgen05 3xTF32 version that makes
// this HBM-bound is the next step for this kernel (DESIGN.md section "fusion_proj").
#include "lgcn_common.cuh"

extern "C" int lgcn_fusion_fwd_tc_try(const float *Eid, const float *C, const float *W, const float *b,
                                      int64_t n_items, int32_t d, int32_t c, float *H, cudaStream_t st);
extern "C" int lgcn_fusion_bwd_w_tc_try(const float *Eid, const float *C, const float *H, const float *gH,
                                        int64_t n_items, int32_t d, int32_t c, float *gW, float *gb,
                                        cudaStream_t st);
extern "C" int lgcn_fusion_bwd_eid_tc_try(const float *W, const float *H, const float *gH, int64_t n_items,
                                          int32_t d, int32_t c, float *gEid, cudaStream_t st);
static int g_fusion_force_simt = 0;

namespace lgcn {

constexpr int FM = 64;   // items per CTA tile
constexpr int FK = 32;   // reduction chunk
constexpr int kFusThreads = 256;

__device__ __forceinline__ float leaky(float h) { return h > 0.0f ? h : 0.01f * h; }

// ---------------------------------------------------------------------------------------
// forward: H[i, o] = leaky(b[o] + sum_k X[i,k] W[o,k]),  X = [Eid | C]
// CTA: 64 items x D outputs; thread (ty = tid/16 -> 4 items, tx = tid%16 -> D/16 outputs)
// ---------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kFusThreads)
fusion_fwd_kernel(const float *__restrict__ Eid, const float *__restrict__ C,
                  const float *__restrict__ W, const float *__restrict__ b, int64_t n_items, int c,
                  float *__restrict__ H) {
    constexpr int NO = D / 16;
    __shared__ float As[FK][FM + 4];   // As[j][item]
    __shared__ float Bs[FK][D + 4];    // Bs[j][out]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t i0 = (int64_t)blockIdx.x * FM;
    const int kin = D + c;
    float acc[4][NO];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int o = 0; o < NO; ++o) acc[a][o] = 0.f;

    for (int k0 = 0; k0 < kin; k0 += FK) {
        for (int idx = tid; idx < FM * (FK / 4); idx += kFusThreads) {
            const int i = idx % FM, jq = idx / FM;
            const int k = k0 + jq * 4;
            const int64_t item = i0 + i;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (item < n_items && k < kin)
                v = (k < D) ? ld_nc_f4(Eid + (size_t)item * D + k)
                            : ld_stream_f4(C + (size_t)item * c + (k - D));
            As[jq * 4 + 0][i] = v.x; As[jq * 4 + 1][i] = v.y;
            As[jq * 4 + 2][i] = v.z; As[jq * 4 + 3][i] = v.w;
        }
        for (int idx = tid; idx < D * (FK / 4); idx += kFusThreads) {
            const int o = idx % D, jq = idx / D;
            const int k = k0 + jq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < kin) v = ld_nc_f4(W + (size_t)o * kin + k);
            Bs[jq * 4 + 0][o] = v.x; Bs[jq * 4 + 1][o] = v.y;
            Bs[jq * 4 + 2][o] = v.z; Bs[jq * 4 + 3][o] = v.w;
        }
        __syncthreads();
#pragma unroll 8
        for (int j = 0; j < FK; ++j) {
            float av[4], bv[NO];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = As[j][ty * 4 + a];
#pragma unroll
            for (int o = 0; o < NO; ++o) bv[o] = Bs[j][tx * NO + o];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int o = 0; o < NO; ++o) acc[a][o] = fmaf(av[a], bv[o], acc[a][o]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t item = i0 + ty * 4 + a;
        if (item >= n_items) continue;
#pragma unroll
        for (int o = 0; o < NO; ++o) {
            const int oo = tx * NO + o;
            H[(size_t)item * D + oo] = leaky(acc[a][o] + __ldg(b + oo));
        }
    }
}

// ---------------------------------------------------------------------------------------
// backward 1: gEid[i,k] = sum_o gpre[i,o] W[o,k]   (k < D),  gpre = gH * leaky'(H)
// ---------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kFusThreads)
fusion_bwd_eid_kernel(const float *__restrict__ W, const float *__restrict__ H,
                      const float *__restrict__ gH, int64_t n_items, int c,
                      float *__restrict__ gEid) {
    constexpr int NO = D / 16;
    __shared__ float As[FK][FM + 4];   // As[o][item] = gpre
    __shared__ float Bs[FK][D + 4];    // Bs[o][k]   = W[o][k]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t i0 = (int64_t)blockIdx.x * FM;
    const int kin = D + c;
    float acc[4][NO];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int o = 0; o < NO; ++o) acc[a][o] = 0.f;
    for (int o0 = 0; o0 < D; o0 += FK) {
        for (int idx = tid; idx < FM * (FK / 4); idx += kFusThreads) {
            const int i = idx % FM, jq = idx / FM;
            const int o = o0 + jq * 4;
            const int64_t item = i0 + i;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (item < n_items && o < D) {
                g = ld_stream_f4(gH + (size_t)item * D + o);
                const float4 h = ld_stream_f4(H + (size_t)item * D + o);
                g.x *= h.x > 0.f ? 1.f : 0.01f; g.y *= h.y > 0.f ? 1.f : 0.01f;
                g.z *= h.z > 0.f ? 1.f : 0.01f; g.w *= h.w > 0.f ? 1.f : 0.01f;
            }
            As[jq * 4 + 0][i] = g.x; As[jq * 4 + 1][i] = g.y;
            As[jq * 4 + 2][i] = g.z; As[jq * 4 + 3][i] = g.w;
        }
        for (int idx = tid; idx < FK * (D / 4); idx += kFusThreads) {
            const int kq = idx % (D / 4), oj = idx / (D / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (o0 + oj < D) v = ld_nc_f4(W + (size_t)(o0 + oj) * kin + kq * 4);
            *reinterpret_cast<float4 *>(&Bs[oj][kq * 4]) = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int j = 0; j < FK; ++j) {
            float av[4], bv[NO];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = As[j][ty * 4 + a];
#pragma unroll
            for (int o = 0; o < NO; ++o) bv[o] = Bs[j][tx * NO + o];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int o = 0; o < NO; ++o) acc[a][o] = fmaf(av[a], bv[o], acc[a][o]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t item = i0 + ty * 4 + a;
        if (item >= n_items) continue;
#pragma unroll
        for (int o = 0; o < NO; ++o) gEid[(size_t)item * D + tx * NO + o] = acc[a][o];
    }
}

// ---------------------------------------------------------------------------------------
// backward 2: gW[o,k] += sum_i gpre[i,o] X[i,k];  gb[o] += sum_i gpre[i,o]
// CTA: (k tile of 64 columns, chunk of items); partial tile accumulated with float atomics.
// thread (ty = tid/16 -> D/16 outputs, tx = tid%16 -> 4 columns)
// ---------------------------------------------------------------------------------------
constexpr int GW_KT = 64;
constexpr int GW_IT = 32;

template <int D>
__global__ void __launch_bounds__(kFusThreads)
fusion_bwd_w_kernel(const float *__restrict__ Eid, const float *__restrict__ C,
                    const float *__restrict__ H, const float *__restrict__ gH, int64_t n_items,
                    int c, int64_t items_per_cta, float *__restrict__ gW, float *__restrict__ gb) {
    constexpr int NO = D / 16;
    __shared__ float As[GW_IT][D + 4];       // gpre[i][o]
    __shared__ float Bs[GW_IT][GW_KT + 4];   // X[i][k]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int kin = D + c;
    const int k0 = blockIdx.x * GW_KT;
    const int64_t ibeg = (int64_t)blockIdx.y * items_per_cta;
    const int64_t iend = min(n_items, ibeg + items_per_cta);
    float acc[NO][4];
    float bsum[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        bsum[o] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[o][q] = 0.f;
    }
    for (int64_t i0 = ibeg; i0 < iend; i0 += GW_IT) {
        for (int idx = tid; idx < GW_IT * (D / 4); idx += kFusThreads) {
            const int oq = idx % (D / 4), i = idx / (D / 4);
            const int64_t item = i0 + i;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (item < iend) {
                g = ld_stream_f4(gH + (size_t)item * D + oq * 4);
                const float4 h = ld_stream_f4(H + (size_t)item * D + oq * 4);
                g.x *= h.x > 0.f ? 1.f : 0.01f; g.y *= h.y > 0.f ? 1.f : 0.01f;
                g.z *= h.z > 0.f ? 1.f : 0.01f; g.w *= h.w > 0.f ? 1.f : 0.01f;
            }
            *reinterpret_cast<float4 *>(&As[i][oq * 4]) = g;
        }
        for (int idx = tid; idx < GW_IT * (GW_KT / 4); idx += kFusThreads) {
            const int kq = idx % (GW_KT / 4), i = idx / (GW_KT / 4);
            const int64_t item = i0 + i;
            const int k = k0 + kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (item < iend && k < kin)
                v = (k < D) ? ld_nc_f4(Eid + (size_t)item * D + k)
                            : ld_stream_f4(C + (size_t)item * c + (k - D));
            *reinterpret_cast<float4 *>(&Bs[i][kq * 4]) = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int i = 0; i < GW_IT; ++i) {
            float av[NO];
#pragma unroll
            for (int o = 0; o < NO; ++o) av[o] = As[i][ty * NO + o];
            const float4 bq = *reinterpret_cast<const float4 *>(&Bs[i][tx * 4]);
            const float bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                bsum[o] += av[o];
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[o][q] = fmaf(av[o], bv[q], acc[o][q]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        const int oo = ty * NO + o;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + tx * 4 + q;
            if (k < kin) atomicAdd(gW + (size_t)oo * kin + k, acc[o][q]);
        }
        if (blockIdx.x == 0 && tx == 0) atomicAdd(gb + oo, bsum[o]);
    }
}

template <int D>
static int fusion_fwd_launch(const float *Eid, const float *C, const float *W, const float *b,
                             int64_t n_items, int c, float *H, cudaStream_t st) {
    const int64_t blocks = (n_items + FM - 1) / FM;
    fusion_fwd_kernel<D><<<(unsigned)blocks, kFusThreads, 0, st>>>(Eid, C, W, b, n_items, c, H);
    LGCN_LAUNCH_CHECK();
    return 0;
}

template <int D>
static int fusion_bwd_launch(const float *Eid, const float *C, const float *W, const float *H,
                             const float *gH, int64_t n_items, int c, float *gEid, float *gW,
                             float *gb, cudaStream_t st) {
    int eid_rc = -100;
    if (!g_fusion_force_simt) eid_rc = lgcn_fusion_bwd_eid_tc_try(W, H, gH, n_items, D, c, gEid, st);
    if (eid_rc > 0 || (eid_rc < 0 && eid_rc != -100)) return eid_rc;
    if (eid_rc == -100) {
        const int64_t blocks = (n_items + FM - 1) / FM;
        fusion_bwd_eid_kernel<D><<<(unsigned)blocks, kFusThreads, 0, st>>>(W, H, gH, n_items, c, gEid);
        LGCN_LAUNCH_CHECK();
    }
    if (!g_fusion_force_simt) {             // tensor-core 3xTF32 path (lgcn_fusion_tc.cu)
        const int rc = lgcn_fusion_bwd_w_tc_try(Eid, C, H, gH, n_items, D, c, gW, gb, st);
        if (rc != -100) return rc;
    }
    const int kin = D + c;
    const int ktiles = (kin + GW_KT - 1) / GW_KT;
    // enough CTAs to fill the chip a few times, few enough to keep the atomic traffic small
    int64_t chunks = (148 * 8 + ktiles - 1) / ktiles;
    int64_t per = (n_items + chunks - 1) / chunks;
    per = ((per + GW_IT - 1) / GW_IT) * GW_IT;
    if (per < GW_IT) per = GW_IT;
    chunks = (n_items + per - 1) / per;
    if (chunks > 65535) { per = ((n_items / 65535 + GW_IT) / GW_IT) * GW_IT; chunks = (n_items + per - 1) / per; }
    dim3 grid((unsigned)ktiles, (unsigned)chunks);
    fusion_bwd_w_kernel<D><<<grid, kFusThreads, 0, st>>>(Eid, C, H, gH, n_items, c, per, gW, gb);
    LGCN_LAUNCH_CHECK();
    return 0;
}

}  // namespace lgcn

// test hook: 1 = always use the fp32 SIMT kernels (the tensor-core path is the default)
extern "C" LGCN_API void lgcn_fusion_force_simt(int on) { g_fusion_force_simt = on; }

extern "C" int lgcn_fusion_proj_fwd(const float *Eid, const float *C, const float *W,
                                    const float *b, int64_t n_items, int32_t d, int32_t c,
                                    float *H, lgcn_stream_t stream) {
    using namespace lgcn;
    if (!dim_supported(d)) return LGCN_E_BAD_DIM;
    if (c <= 0 || c % 4 != 0) return LGCN_E_BAD_DIM;
    if (n_items < 0 || !Eid || !C || !W || !b || !H) return LGCN_E_BAD_ARG;
    if (n_items == 0) return 0;
    if ((n_items + FM - 1) / FM > 0x7fffffffLL) return LGCN_E_TOO_LARGE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!(g_fusion_force_simt)) {           // tensor-core 3xTF32 path (lgcn_fusion_tc.cu)
        const int rc = lgcn_fusion_fwd_tc_try(Eid, C, W, b, n_items, d, c, H, st);
        if (rc != -100) return rc;
    }
    switch (d) {
        case 16:  return fusion_fwd_launch<16>(Eid, C, W, b, n_items, c, H, st);
        case 32:  return fusion_fwd_launch<32>(Eid, C, W, b, n_items, c, H, st);
        case 64:  return fusion_fwd_launch<64>(Eid, C, W, b, n_items, c, H, st);
        case 128: return fusion_fwd_launch<128>(Eid, C, W, b, n_items, c, H, st);
        case 256: return fusion_fwd_launch<256>(Eid, C, W, b, n_items, c, H, st);
    }
    return LGCN_E_BAD_DIM;
}

extern "C" int lgcn_fusion_proj_bwd(const float *Eid, const float *C, const float *W,
                                    const float *H, const float *gH, int64_t n_items, int32_t d,
                                    int32_t c, float *gEid, float *gW, float *gb,
                                    lgcn_stream_t stream) {
    using namespace lgcn;
    if (!dim_supported(d)) return LGCN_E_BAD_DIM;
    if (c <= 0 || c % 4 != 0) return LGCN_E_BAD_DIM;
    if (n_items < 0 || !Eid || !C || !W || !H || !gH || !gEid || !gW || !gb) return LGCN_E_BAD_ARG;
    if (n_items == 0) return 0;
    if ((n_items + FM - 1) / FM > 0x7fffffffLL) return LGCN_E_TOO_LARGE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (d) {
        case 16:  return fusion_bwd_launch<16>(Eid, C, W, H, gH, n_items, c, gEid, gW, gb, st);
        case 32:  return fusion_bwd_launch<32>(Eid, C, W, H, gH, n_items, c, gEid, gW, gb, st);
        case 64:  return fusion_bwd_launch<64>(Eid, C, W, H, gH, n_items, c, gEid, gW, gb, st);
        case 128: return fusion_bwd_launch<128>(Eid, C, W, H, gH, n_items, c, gEid, gW, gb, st);
        case 256: return fusion_bwd_launch<256>(Eid, C, W, H, gH, n_items, c, gEid, gW, gb, st);
    }
    return LGCN_E_BAD_DIM;
}
