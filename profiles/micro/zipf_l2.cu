// Microbenchmark: does the B200 L2 retain the hot rows of a Zipf-distributed
// 512-byte row gather, and what takes them out?  Models the USER-side half of one Amazon-shape
// lgcn_spmm launch: 29.5 M gathers of item rows (4.4 M x 512 B = 2.25 GB table, popularity ~
// rank^-0.8, hot items scattered over the table by a random permutation), one 512-byte output row
// stored per 3 gathers (the Y stream), indices streamed.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zipf_l2 zipf_l2.cu && ./zipf_l2
//
// Variants (each timed with CUDA events, best of 3 after a warm-up launch):
//   uniform        uniform random rows: the pure-DRAM gather rate, no reuse to find
//   hotonly:<MB>   every gather falls into the <MB> hottest rows: the L2-hit gather rate as a
//                  function of footprint = the EFFECTIVE L2 capacity for scattered rows
//   zipf           Zipf(0.8), default policies, with / without the store stream
//   zipf +policy   evict_first stores; evict_last on the <H> hottest rows; evict_first /
//                  no-allocate on the cold rows
//   compact        hot rows relocated to a contiguous block at the head of the table
// "eff GB/s" = gathered bytes / time; anything above the uniform rate is L2 hits.  ncu
// (dram__bytes_read.sum, lts__t_sector_hit_rate.pct) on the same binary gives the hit rates.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kHot = 0x80000000;   // bit 31: hot class
constexpr int kMask = 0x7fffffff;

__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_normal() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_unchanged() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(p)); return p; }

__device__ __forceinline__ float4 ld_hint(const float4 *p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_hint(float4 *p, const float4 &v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// HOTP: 0 normal, 1 evict_last.  COLDP: 0 normal, 1 evict_first, 2 evict_unchanged.
// STORE: 0 none, 1 default policy, 2 evict_first, 3 st.global.cs (streaming)
template <int HOTP, int COLDP, int STORE>
__global__ void __launch_bounds__(128) gather_kernel(const float4 *__restrict__ tab, const int *__restrict__ idx,
                                                     long long n_idx, float4 *__restrict__ out, int per_warp) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long base = warp * per_warp;
    const uint64_t pf = pol_first();
    const uint64_t ph = HOTP == 1 ? pol_last() : pol_normal();
    const uint64_t pc = COLDP == 1 ? pf : (COLDP == 2 ? pol_unchanged() : pol_normal());
    float4 acc = make_float4(0, 0, 0, 0);
    long long orow = base / 3;
    int since = 0;
    for (int i = 0; i < per_warp; i += 8) {
        int r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long e = base + i + u;
            int v = 0;
            if (e < n_idx) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(idx + e), "l"(pf));
            r[u] = v;
        }
        float4 x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            x[u] = ld_hint(tab + (size_t)(r[u] & kMask) * 32 + lane, r[u] < 0 ? ph : pc);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w;
            if (STORE != 0 && ++since == 3) {
                since = 0;
                float4 *dst = out + (size_t)orow * 32 + lane;
                if (STORE == 1) *dst = acc;
                else if (STORE == 2) st_hint(dst, acc, pf);
                else __stcs(dst, acc);
                ++orow;
                acc = make_float4(0, 0, 0, 0);
            }
        }
    }
    if (STORE == 0 && acc.x == 12345.678f) out[warp * 32 + lane] = acc;   // keep the loads alive
}

// idx[e] = row of rank[e] (+ hot bit): scattered = perm[rank]; compact = rank itself for rank < H
__global__ void build_idx(const int *rank, const int *perm, long long n, int n_hot, int compact, int *idx) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int rk = rank[e];
    int row = perm[rk];
    if (compact) row = rk;          // rows sorted by popularity: the hot set is one contiguous block
    idx[e] = row | (rk < n_hot ? kHot : 0);
}

template <int HOTP, int COLDP, int STORE>
static float run(const float4 *tab, const int *idx, long long n_idx, float4 *out) {
    const int per_warp = 24;
    const long long warps = (n_idx + per_warp - 1) / per_warp;
    const int blocks = (int)((warps + 3) / 4);
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    gather_kernel<HOTP, COLDP, STORE><<<blocks, 128>>>(tab, idx, n_idx, out, per_warp);
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(a));
        gather_kernel<HOTP, COLDP, STORE><<<blocks, 128>>>(tab, idx, n_idx, out, per_warp);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char **argv) {
    const long long rows = 4400000, n_idx = 29500000;
    const double alpha = 0.8;
    const bool quick = argc > 1 && !strcmp(argv[1], "quick");   // the subset profiled under ncu
    float4 *tab, *out; int *idx, *rank_d, *perm_d;
    CK(cudaMalloc(&tab, rows * 512)); CK(cudaMemset(tab, 0, rows * 512));
    CK(cudaMalloc(&out, (n_idx / 3 + 64) * 512));
    CK(cudaMalloc(&idx, n_idx * 4)); CK(cudaMalloc(&rank_d, n_idx * 4)); CK(cudaMalloc(&perm_d, rows * 4));
    // host: Zipf ranks by inverse CDF, random rank -> row permutation
    std::vector<double> cdf(rows);
    double s = 0;
    for (long long r = 0; r < rows; ++r) { s += std::pow((double)(r + 1), -alpha); cdf[r] = s; }
    std::mt19937_64 rng(1234);
    std::uniform_real_distribution<double> uni(0.0, 1.0);
    std::vector<int> rank(n_idx), perm(rows);
    for (long long e = 0; e < n_idx; ++e)
        rank[e] = (int)(std::upper_bound(cdf.begin(), cdf.end(), uni(rng) * s) - cdf.begin());
    for (long long r = 0; r < rows; ++r) perm[r] = (int)r;
    std::shuffle(perm.begin(), perm.end(), rng);
    CK(cudaMemcpy(perm_d, perm.data(), rows * 4, cudaMemcpyHostToDevice));
    const double gb = (double)n_idx * 512 / 1e9;
    auto upload = [&](const std::vector<int> &rk) { CK(cudaMemcpy(rank_d, rk.data(), n_idx * 4, cudaMemcpyHostToDevice)); };
    auto mk = [&](int n_hot, int compact) { build_idx<<<(unsigned)((n_idx + 255) / 256), 256>>>(rank_d, perm_d, n_idx, n_hot, compact, idx); CK(cudaDeviceSynchronize()); };
    auto rep = [&](const char *name, float ms) { printf("%-58s %7.3f ms  eff %6.0f GB/s\n", name, ms, gb / ms * 1e3); fflush(stdout); };
    char nm[128];

    // 1. uniform: pure DRAM gather rate
    {
        std::vector<int> u(n_idx);
        for (long long e = 0; e < n_idx; ++e) u[e] = (int)(rng() % (unsigned long long)rows);
        upload(u); mk(0, 0);
        rep("uniform, no stores", run<0, 0, 0>(tab, idx, n_idx, out));
        rep("uniform, stores evict_first", run<0, 0, 2>(tab, idx, n_idx, out));
    }
    // 2. hot-only footprints: effective L2 capacity for scattered 512-byte rows
    if (!quick)
    for (int mb : {16, 32, 48, 64, 80, 96, 112, 126, 160}) {
        const long long hot_rows = (long long)mb * (1 << 20) / 512;
        std::vector<int> u(n_idx);
        for (long long e = 0; e < n_idx; ++e) u[e] = (int)(rng() % (unsigned long long)hot_rows);
        upload(u); mk(0, 0);
        snprintf(nm, sizeof nm, "hotonly %3d MB uniform inside, no stores", mb);
        rep(nm, run<0, 0, 0>(tab, idx, n_idx, out));
        snprintf(nm, sizeof nm, "hotonly %3d MB uniform inside, stores evict_first", mb);
        rep(nm, run<0, 0, 2>(tab, idx, n_idx, out));
    }
    // 3. Zipf
    upload(rank);
    mk(0, 0);
    rep("zipf, no stores", run<0, 0, 0>(tab, idx, n_idx, out));
    rep("zipf, stores default", run<0, 0, 1>(tab, idx, n_idx, out));
    rep("zipf, stores evict_first", run<0, 0, 2>(tab, idx, n_idx, out));
    if (!quick) rep("zipf, stores st.cs", run<0, 0, 3>(tab, idx, n_idx, out));
    rep("zipf, cold evict_first (all), stores evict_first", run<0, 1, 2>(tab, idx, n_idx, out));
    for (int mb : {16, 32, 64, 96}) {
        if (quick && mb != 64) continue;
        const int n_hot = (int)((long long)mb * (1 << 20) / 512);
        mk(n_hot, 0);
        snprintf(nm, sizeof nm, "zipf hot %3d MB evict_last, no stores", mb);
        rep(nm, run<1, 0, 0>(tab, idx, n_idx, out));
        snprintf(nm, sizeof nm, "zipf hot %3d MB evict_last, stores evict_first", mb);
        rep(nm, run<1, 0, 2>(tab, idx, n_idx, out));
        snprintf(nm, sizeof nm, "zipf hot %3d MB evict_last, cold evict_first, st first", mb);
        rep(nm, run<1, 1, 2>(tab, idx, n_idx, out));
        snprintf(nm, sizeof nm, "zipf hot %3d MB normal,     cold evict_first, st first", mb);
        rep(nm, run<0, 1, 2>(tab, idx, n_idx, out));
        if (!quick) {
            snprintf(nm, sizeof nm, "zipf hot %3d MB evict_last, cold unchanged,  st first", mb);
            rep(nm, run<1, 2, 2>(tab, idx, n_idx, out));
        }
        mk(n_hot, 1);
        snprintf(nm, sizeof nm, "zipf COMPACT hot %3d MB evict_last, cold first, st first", mb);
        rep(nm, run<1, 1, 2>(tab, idx, n_idx, out));
        if (!quick) {
            snprintf(nm, sizeof nm, "zipf COMPACT hot %3d MB normal, stores evict_first", mb);
            rep(nm, run<0, 0, 2>(tab, idx, n_idx, out));
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
