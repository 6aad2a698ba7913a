// lgcn_bpr.cu -- fused BPR + L2 step and dense Adam (sm_100a).
//
// lgcn_bpr_fused replaces the six row gathers of reference main.py:496-497, bpr_loss_reg
// (reference main.py:366-402) and their autograd backward (index_put accumulate): one warp
// per (user, pos, neg) sample gathers the three propagated rows and the three layer-0 rows,
// forms the two dots with shuffles, evaluates -log(sigmoid(x)+1e-8) and its derivative, and
// scatter-adds the six gradient rows with vector float atomics (red.global.add.v4.f32).
// lgcn_adam replaces torch.optim.Adam.step (reference main.py:469,526).
#include "lgcn_common.cuh"

namespace lgcn {

constexpr int kBprThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// D/4 lanes of the warp are active (D <= 128), or every lane owns D/128 float4 (D == 256).
// PHASE 0: fused single-GPU step.  PHASE 1: partial dots / norms only (feature-sharded tables:
// every rank owns d/P columns) -> dots[0:bs]=<u,p>, dots[bs:2bs]=<u,n>, dots[2bs:3bs]=|.|^2.
// PHASE 2: dots already summed over ranks; form the loss terms and scatter this rank's columns.
template <int D, int PHASE>
__global__ void __launch_bounds__(kBprThreads)
bpr_kernel(const float *__restrict__ F, const float *__restrict__ P,
           const int64_t *__restrict__ users, const int64_t *__restrict__ pos,
           const int64_t *__restrict__ neg, int64_t bs, int64_t item_offset, float lam,
           float grad_scale, int flags, float *__restrict__ dots, float *__restrict__ sample_ws,
           float *__restrict__ gF, float *__restrict__ gP, uint8_t *__restrict__ rowflag) {
    using G = RowGeom<D>;
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * (kBprThreads / 32) + (threadIdx.x >> 5);
    if (s >= bs) return;
    const bool act = lane < G::LANES;
    const int64_t ru = users[s], rp = item_offset + pos[s], rn = item_offset + neg[s];
    float4 fu[G::VEC], fp[G::VEC], fn[G::VEC], eu[G::VEC], ep[G::VEC], en[G::VEC];
    float ps = 0.f, ns = 0.f, reg = 0.f;
    if (act) {
#pragma unroll
        for (int v = 0; v < G::VEC; ++v) {
            const int o = lane * 4 + v * G::LANES * 4;
            fu[v] = ld_nc_f4(F + ru * D + o);
            fp[v] = ld_nc_f4(F + rp * D + o);
            fn[v] = ld_nc_f4(F + rn * D + o);
            eu[v] = ld_nc_f4(P + ru * D + o);
            ep[v] = ld_nc_f4(P + rp * D + o);
            en[v] = ld_nc_f4(P + rn * D + o);
        }
        if (PHASE != 2) {
#pragma unroll
            for (int v = 0; v < G::VEC; ++v) {
                ps += fu[v].x * fp[v].x + fu[v].y * fp[v].y + fu[v].z * fp[v].z + fu[v].w * fp[v].w;
                ns += fu[v].x * fn[v].x + fu[v].y * fn[v].y + fu[v].z * fn[v].z + fu[v].w * fn[v].w;
                reg += eu[v].x * eu[v].x + eu[v].y * eu[v].y + eu[v].z * eu[v].z + eu[v].w * eu[v].w;
                reg += ep[v].x * ep[v].x + ep[v].y * ep[v].y + ep[v].z * ep[v].z + ep[v].w * ep[v].w;
                reg += en[v].x * en[v].x + en[v].y * en[v].y + en[v].z * en[v].z + en[v].w * en[v].w;
            }
        }
    }
    if (PHASE != 2) {
        ps = warp_sum(ps);
        ns = warp_sum(ns);
        reg = warp_sum(reg);
    }
    if (PHASE == 1) {
        if (lane == 0) { dots[s] = ps; dots[bs + s] = ns; dots[2 * bs + s] = reg; }
        return;
    }
    if (PHASE == 2) { ps = dots[s]; ns = dots[bs + s]; reg = dots[2 * bs + s]; }
    const float x = ps - ns;                       // main.py:377-379
    const float sg = 1.0f / (1.0f + expf(-x));
    if (lane == 0) {
        sample_ws[s] = -logf(sg + 1e-8f);
        sample_ws[bs + s] = reg;
    }
    if (flags & LGCN_BPR_NO_GRAD) return;
    const float invB = 1.0f / (float)bs;
    // d/dx -log(sigmoid(x)+1e-8) = -sg(1-sg)/(sg+1e-8); mean over the batch
    const float coef = -sg * (1.0f - sg) / (sg + 1e-8f) * invB * grad_scale;
    const float c2 = 2.0f * lam * invB;            // main.py:394-398 (no 1/2)
    if (rowflag && lane == 0) { rowflag[ru] = 1; rowflag[rp] = 1; rowflag[rn] = 1; }
    if (!act) return;
    const bool both = (flags & LGCN_BPR_GP_INCLUDES_GF) != 0;
#pragma unroll
    for (int v = 0; v < G::VEC; ++v) {
        const int o = lane * 4 + v * G::LANES * 4;
        const float4 du = make_float4(coef * (fp[v].x - fn[v].x), coef * (fp[v].y - fn[v].y),
                                      coef * (fp[v].z - fn[v].z), coef * (fp[v].w - fn[v].w));
        const float4 dp = make_float4(coef * fu[v].x, coef * fu[v].y, coef * fu[v].z, coef * fu[v].w);
        const float4 dn = make_float4(-dp.x, -dp.y, -dp.z, -dp.w);
        if (gF) {
            atomicAdd(reinterpret_cast<float4 *>(gF + ru * D + o), du);
            atomicAdd(reinterpret_cast<float4 *>(gF + rp * D + o), dp);
            atomicAdd(reinterpret_cast<float4 *>(gF + rn * D + o), dn);
        }
        if (gP) {
            float4 hu = make_float4(c2 * eu[v].x, c2 * eu[v].y, c2 * eu[v].z, c2 * eu[v].w);
            float4 hp = make_float4(c2 * ep[v].x, c2 * ep[v].y, c2 * ep[v].z, c2 * ep[v].w);
            float4 hn = make_float4(c2 * en[v].x, c2 * en[v].y, c2 * en[v].z, c2 * en[v].w);
            if (both) {
                hu.x += du.x; hu.y += du.y; hu.z += du.z; hu.w += du.w;
                hp.x += dp.x; hp.y += dp.y; hp.z += dp.z; hp.w += dp.w;
                hn.x += dn.x; hn.y += dn.y; hn.z += dn.z; hn.w += dn.w;
            }
            atomicAdd(reinterpret_cast<float4 *>(gP + ru * D + o), hu);
            atomicAdd(reinterpret_cast<float4 *>(gP + rp * D + o), hp);
            atomicAdd(reinterpret_cast<float4 *>(gP + rn * D + o), hn);
        }
    }
}

// loss = mean(sample_ws[0:bs]) + lam * sum(sample_ws[bs:2bs]) / bs, fixed summation order.
__global__ void __launch_bounds__(1024) bpr_reduce_kernel(const float *__restrict__ ws, int64_t bs,
                                                          float lam, float *__restrict__ loss_out) {
    __shared__ double sh_a[32], sh_b[32];
    double a = 0.0, b = 0.0;
    for (int64_t i = threadIdx.x; i < bs; i += blockDim.x) {
        a += (double)ws[i];
        b += (double)ws[bs + i];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    if ((threadIdx.x & 31) == 0) { sh_a[threadIdx.x >> 5] = a; sh_b[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ta += sh_a[w]; tb += sh_b[w]; }
        loss_out[0] = (float)(ta / (double)bs + (double)lam * tb / (double)bs);
    }
}

template <int D>
__global__ void __launch_bounds__(kBprThreads)
zero_rows_kernel(float *__restrict__ t0, float *__restrict__ t1, uint8_t *__restrict__ rowflag,
                 const int64_t *__restrict__ users, const int64_t *__restrict__ pos,
                 const int64_t *__restrict__ neg, int64_t bs, int64_t item_offset) {
    using G = RowGeom<D>;
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * (kBprThreads / 32) + (threadIdx.x >> 5);
    if (s >= bs || lane >= G::LANES) return;
    const int64_t rows[3] = {users[s], item_offset + pos[s], item_offset + neg[s]};
    if (rowflag && lane == 0) { rowflag[rows[0]] = 0; rowflag[rows[1]] = 0; rowflag[rows[2]] = 0; }
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int v = 0; v < G::VEC; ++v) {
            const size_t o = (size_t)rows[k] * D + lane * 4 + v * G::LANES * 4;
            if (t0) st_f4(t0 + o, z);
            if (t1) st_f4(t1 + o, z);
        }
}

__global__ void adam_tick_kernel(int64_t *step, float *scalars, float lr, float beta1, float beta2) {
    const int64_t t = step[0] + 1;
    step[0] = t;
    const double bc1 = 1.0 - pow((double)beta1, (double)t);
    const double bc2 = 1.0 - pow((double)beta2, (double)t);
    scalars[0] = (float)((double)lr / bc1);
    scalars[1] = (float)sqrt(bc2);
}

__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ p, const float *__restrict__ g0, const float *__restrict__ g1,
            float *__restrict__ m, float *__restrict__ v, int64_t n4, int64_t n,
            const float *__restrict__ scalars, float beta1, float beta2, float eps) {
    const float ss = __ldg(scalars), bs = __ldg(scalars + 1);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 g = ld_stream_f4(g0 + i * 4);
        if (g1) { const float4 t = ld_stream_f4(g1 + i * 4); add4(g, t); }
        float4 pp = *reinterpret_cast<float4 *>(p + i * 4);
        float4 mm = *reinterpret_cast<float4 *>(m + i * 4);
        float4 vv = *reinterpret_cast<float4 *>(v + i * 4);
        adam4(pp, mm, vv, g, ss, bs, beta1, beta2, eps);
        st_f4(p + i * 4, pp);
        st_f4(m + i * 4, mm);
        st_f4(v + i * 4, vv);
    }
    // tail (n not a multiple of 4): handled by the first threads
    const int64_t tail = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tail < n) {
        float g = g0[tail] + (g1 ? g1[tail] : 0.0f);
        adam_elem(p[tail], m[tail], v[tail], g, ss, bs, beta1, beta2, eps);
    }
}

}  // namespace lgcn

template <int PHASE>
static int bpr_launch(const float *F, const float *P, const int64_t *users, const int64_t *pos,
                      const int64_t *neg, int64_t bs, int32_t d, int64_t item_offset, float lam,
                      float grad_scale, int32_t flags, float *dots, float *sample_ws, float *gF,
                      float *gP, uint8_t *rowflag, cudaStream_t st) {
    using namespace lgcn;
    const unsigned grid = (unsigned)((bs + kBprThreads / 32 - 1) / (kBprThreads / 32));
#define LGCN_BPR_CASE(DD)                                                                          \
    case DD:                                                                                       \
        bpr_kernel<DD, PHASE><<<grid, kBprThreads, 0, st>>>(F, P, users, pos, neg, bs, item_offset, \
                                                            lam, grad_scale, flags, dots, sample_ws, \
                                                            gF, gP, rowflag);                      \
        break;
    switch (d) {
        LGCN_BPR_CASE(16) LGCN_BPR_CASE(32) LGCN_BPR_CASE(64) LGCN_BPR_CASE(128) LGCN_BPR_CASE(256)
        default: return LGCN_E_BAD_DIM;
    }
#undef LGCN_BPR_CASE
    LGCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int lgcn_bpr_fused(const float *F, const float *P, const int64_t *users,
                              const int64_t *pos, const int64_t *neg, int64_t bs, int32_t d,
                              int64_t item_offset, float lam, float grad_scale, int32_t flags,
                              float *sample_ws, float *loss_out, float *gF, float *gP,
                              uint8_t *rowflag, lgcn_stream_t stream) {
    using namespace lgcn;
    if (!dim_supported(d)) return LGCN_E_BAD_DIM;
    if (bs <= 0 || !F || !P || !users || !pos || !neg || !sample_ws || !loss_out) return LGCN_E_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = bpr_launch<0>(F, P, users, pos, neg, bs, d, item_offset, lam, grad_scale, flags, nullptr,
                           sample_ws, gF, gP, rowflag, st);
    if (rc) return rc;
    bpr_reduce_kernel<<<1, 1024, 0, st>>>(sample_ws, bs, lam, loss_out);
    LGCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int lgcn_bpr_partial(const float *F, const float *P, const int64_t *users,
                                const int64_t *pos, const int64_t *neg, int64_t bs, int32_t d,
                                int64_t item_offset, float *dots, lgcn_stream_t stream) {
    using namespace lgcn;
    if (!dim_supported(d)) return LGCN_E_BAD_DIM;
    if (bs <= 0 || !F || !P || !users || !pos || !neg || !dots) return LGCN_E_BAD_ARG;
    return bpr_launch<1>(F, P, users, pos, neg, bs, d, item_offset, 0.f, 0.f, 0, dots, nullptr,
                         nullptr, nullptr, nullptr, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int lgcn_bpr_apply(const float *F, const float *P, const int64_t *users,
                              const int64_t *pos, const int64_t *neg, int64_t bs, int32_t d,
                              int64_t item_offset, float lam, float grad_scale, int32_t flags,
                              const float *dots, float *sample_ws, float *loss_out, float *gF,
                              float *gP, uint8_t *rowflag, lgcn_stream_t stream) {
    using namespace lgcn;
    if (!dim_supported(d)) return LGCN_E_BAD_DIM;
    if (bs <= 0 || !F || !P || !users || !pos || !neg || !dots || !sample_ws || !loss_out)
        return LGCN_E_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = bpr_launch<2>(F, P, users, pos, neg, bs, d, item_offset, lam, grad_scale, flags,
                           const_cast<float *>(dots), sample_ws, gF, gP, rowflag, st);
    if (rc) return rc;
    bpr_reduce_kernel<<<1, 1024, 0, st>>>(sample_ws, bs, lam, loss_out);
    LGCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int lgcn_zero_rows(float *t0, float *t1, uint8_t *rowflag, const int64_t *users,
                              const int64_t *pos, const int64_t *neg, int64_t bs, int32_t d,
                              int64_t item_offset, lgcn_stream_t stream) {
    using namespace lgcn;
    if (!dim_supported(d)) return LGCN_E_BAD_DIM;
    if (bs <= 0 || !users || !pos || !neg) return LGCN_E_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)((bs + kBprThreads / 32 - 1) / (kBprThreads / 32));
#define LGCN_ZR_CASE(DD)                                                                          \
    case DD:                                                                                      \
        zero_rows_kernel<DD><<<grid, kBprThreads, 0, st>>>(t0, t1, rowflag, users, pos, neg, bs,  \
                                                           item_offset);                          \
        break;
    switch (d) { LGCN_ZR_CASE(16) LGCN_ZR_CASE(32) LGCN_ZR_CASE(64) LGCN_ZR_CASE(128) LGCN_ZR_CASE(256) }
#undef LGCN_ZR_CASE
    LGCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int lgcn_adam_tick(int64_t *step_dev, float *adam_scalars, float lr, float beta1,
                              float beta2, lgcn_stream_t stream) {
    if (!step_dev || !adam_scalars) return LGCN_E_BAD_ARG;
    lgcn::adam_tick_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(step_dev, adam_scalars,
                                                                              lr, beta1, beta2);
    LGCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int lgcn_adam(float *p, const float *g0, const float *g1, float *m, float *v, int64_t n,
                         const float *adam_scalars, float beta1, float beta2, float eps,
                         lgcn_stream_t stream) {
    if (n < 0 || !p || !g0 || !m || !v || !adam_scalars) return LGCN_E_BAD_ARG;
    if (n == 0) return 0;
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    lgcn::adam_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        p, g0, g1, m, v, n4, n, adam_scalars, beta1, beta2, eps);
    LGCN_LAUNCH_CHECK();
    return 0;
}
