// Shared device helpers for the LightGCN sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "lgcn.h"

#define LGCN_LAUNCH_CHECK()                         \
    do {                                            \
        cudaError_t e__ = cudaGetLastError();       \
        if (e__ != cudaSuccess) return (int)e__;    \
    } while (0)

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute is per DEVICE (and
// context), so the "done" cache is a bit per device ordinal; a race between host threads only
// repeats the idempotent call.  Use inside a function returning int (0 = ok).  Wrap a kernel name
// that contains commas in parentheses.
#define LGCN_OPT_IN_SMEM(kernel, bytes)                                                         \
    do {                                                                                        \
        static std::atomic<unsigned long long> done__{0ull};                                    \
        int dev__ = 0;                                                                          \
        cudaError_t e__ = cudaGetDevice(&dev__);                                                \
        if (e__ != cudaSuccess) return (int)e__;                                                \
        const unsigned long long bit__ = 1ull << (dev__ & 63);                                  \
        if (!(done__.load(std::memory_order_acquire) & bit__)) {                                \
            e__ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                       (int)(bytes));                                           \
            if (e__ != cudaSuccess) return (int)e__;                                            \
            done__.fetch_or(bit__, std::memory_order_release);                                  \
        }                                                                                       \
    } while (0)

namespace lgcn {

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline bool dim_supported(int d) {
    return d == 16 || d == 32 || d == 64 || d == 128 || d == 256;
}

// Geometry of one table row spread over a sub-warp group of lanes, float4 per lane.
template <int D>
struct RowGeom {
    static constexpr int VEC = (D > 128) ? D / 128 : 1;  // float4 per lane
    static constexpr int LANES = D / (4 * VEC);          // lanes that own one row
    static constexpr int GROUPS = 32 / LANES;            // rows per warp
    static_assert(LANES >= 4 && LANES <= 32 && (LANES & (LANES - 1)) == 0, "bad dim");
};

__device__ __forceinline__ float4 ld_nc_f4(const float *p) {
    return __ldg(reinterpret_cast<const float4 *>(p));
}
// L2 eviction policies (createpolicy) and loads / stores that carry them.  Tables are far
// larger than L2 at the Amazon shape: streaming data (entries, outputs, epilogue operands) is
// marked evict_first so it does not displace the gathered rows of high-degree columns, which
// are marked evict_last by the graph plan (bit 31 of the packed column index).
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ld_nc_f4_hint(const float *p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float4 ld_stream_f4_hint(const float *p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ int2 ld_stream_i2_hint(const int2 *p, uint64_t pol) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.s32 {%0,%1}, [%2], %3;"
                 : "=r"(r.x), "=r"(r.y)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_f4_hint(float *p, const float4 &v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
                 : "memory");
}
// streaming (read-once) loads: keep them out of L1
__device__ __forceinline__ float4 ld_stream_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ int ld_stream_i32(const int32_t *p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f32(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_f4(float *p, const float4 &v) {
    *reinterpret_cast<float4 *>(p) = v;
}
__device__ __forceinline__ void fma4(float4 &acc, float w, const float4 &x) {
    acc.x = fmaf(w, x.x, acc.x);
    acc.y = fmaf(w, x.y, acc.y);
    acc.z = fmaf(w, x.z, acc.z);
    acc.w = fmaf(w, x.w, acc.w);
}
__device__ __forceinline__ void add4(float4 &a, const float4 &b) {
    a.x = __fadd_rn(a.x, b.x);
    a.y = __fadd_rn(a.y, b.y);
    a.z = __fadd_rn(a.z, b.z);
    a.w = __fadd_rn(a.w, b.w);
}

// One element of torch.optim.Adam's single-tensor update (reference main.py:469,526):
// m.lerp_(g,1-b1); v.mul_(b2).addcmul_(g,g,1-b2); p -= step*(m/(sqrt(v)/bc2s+eps)).
__device__ __forceinline__ void adam_elem(float &p, float &m, float &v, float g, float step_size,
                                          float bc2_sqrt, float beta1, float beta2, float eps) {
    m = __fmaf_rn(__fsub_rn(g, m), 1.0f - beta1, m);
    v = __fmaf_rn((1.0f - beta2) * g, g, v * beta2);
#ifdef LGCN_ADAM_EXACT    // correctly rounded sqrt / divisions (the CPU arithmetic, bit for bit)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);
    p = __fsub_rn(p, step_size * __fdiv_rn(m, denom));
#else
    // fp32 sqrt.approx / div.approx (<= 2 ulp each): the update term step*m/denom is off by
    // < 1e-6 relative, i.e. < 1e-8 of p -- far inside the 1e-5 bar -- and the fused ADAM hop
    // drops from 13.17 to 12.19 ms at the Amazon shape (it is issue/latency bound: ncu
    // profiles/r01_spmm_chunk_adam_amazon_ncu.txt, 39 % issue-active at 26 % occupancy).
    // m and v themselves are computed exactly.
    float sq;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(v));
    const float denom = __fadd_rn(__fdividef(sq, bc2_sqrt), eps);
    p = __fsub_rn(p, step_size * __fdividef(m, denom));
#endif
}

__device__ __forceinline__ void adam4(float4 &p, float4 &m, float4 &v, const float4 &g, float ss,
                                      float bs, float b1, float b2, float eps) {
    adam_elem(p.x, m.x, v.x, g.x, ss, bs, b1, b2, eps);
    adam_elem(p.y, m.y, v.y, g.y, ss, bs, b1, b2, eps);
    adam_elem(p.z, m.z, v.z, g.z, ss, bs, b1, b2, eps);
    adam_elem(p.w, m.w, v.w, g.w, ss, bs, b1, b2, eps);
}

}  // namespace lgcn
