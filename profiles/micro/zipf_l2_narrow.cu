// Microbenchmark: Zipf-distributed gathers of NARROW rows (d/P = 16 -> 64 bytes, d/P = 32 -> 128 bytes),
// the user-side half of one feature-sharded Amazon-shape lgcn_spmm launch at 8 / 4 GPUs.
//
// Question: a 64-byte row occupies HALF a 128-byte L2 line.  When the hot item rows are scattered
// over the table (ids are what the dataset says), each of them holds a whole line's tag -> the
// effective L2 capacity for hot rows is halved.  Does a PHYSICAL row order that packs the hot rows
// next to each other (popularity order; the per-row entry order -- and with it every bit of the SpMM
// result -- can stay as it is) raise the hit rate?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zipf_l2_narrow zipf_l2_narrow.cu && ./zipf_l2_narrow
//
// Variants, for ROWB = 64 and 128 bytes: uniform (pure DRAM rate), zipf scattered, zipf packed
// (rows in popularity order), each with an evict_first store stream of one row per 3 gathers.
// "eff GB/s" = gathered bytes / time.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }

// LANES lanes own one row (float4 each); a warp gathers 32 / LANES rows per instruction.
template <int LANES, bool STORE>
__global__ void __launch_bounds__(128) gather_kernel(const float4 *__restrict__ tab, const int *__restrict__ idx,
                                                     long long n_idx, float4 *__restrict__ out, int per_group) {
    constexpr int GROUPS = 32 / LANES;
    const int lane = threadIdx.x & 31, sub = lane % LANES;
    const long long group = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * GROUPS + lane / LANES;
    const long long base = group * per_group;
    const uint64_t pf = pol_first();
    float4 acc = make_float4(0, 0, 0, 0);
    long long orow = base / 3;
    int since = 0;
    for (int i = 0; i < per_group; i += 8) {
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
        for (int u = 0; u < 8; ++u) x[u] = __ldg(tab + (size_t)r[u] * LANES + sub);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w;
            if (STORE && ++since == 3) {
                since = 0;
                float4 *dst = out + (size_t)orow * LANES + sub;
                asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                             :: "l"(dst), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w), "l"(pf) : "memory");
                ++orow;
                acc = make_float4(0, 0, 0, 0);
            }
        }
    }
    if (!STORE && acc.x == 12345.678f) out[group * LANES + sub] = acc;   // keep the loads alive
}

__global__ void build_idx(const int *rank, const int *perm, long long n, int packed, int *idx) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) idx[e] = packed ? rank[e] : perm[rank[e]];
}

template <int LANES, bool STORE>
static float run(const float4 *tab, const int *idx, long long n_idx, float4 *out) {
    constexpr int GROUPS = 32 / LANES;
    const int per_group = 24;
    const long long groups = (n_idx + per_group - 1) / per_group;
    const int blocks = (int)((groups + 4 * GROUPS - 1) / (4 * GROUPS));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    gather_kernel<LANES, STORE><<<blocks, 128>>>(tab, idx, n_idx, out, per_group);
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(a));
        gather_kernel<LANES, STORE><<<blocks, 128>>>(tab, idx, n_idx, out, per_group);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    return best;
}

template <int LANES>
static void sweep(const float4 *tab, int *idx, const int *rank_d, const int *unif_d, const int *perm_d,
                  long long n_idx, float4 *out) {
    const int rowb = LANES * 16;
    const double gb = (double)n_idx * rowb / 1e9;
    auto mk = [&](const int *src, int packed) { build_idx<<<(unsigned)((n_idx + 255) / 256), 256>>>(src, perm_d, n_idx, packed, idx); CK(cudaDeviceSynchronize()); };
    auto rep = [&](const char *name, float ms) { printf("row %3d B  %-44s %7.3f ms  eff %6.0f GB/s\n", rowb, name, ms, gb / ms * 1e3); fflush(stdout); };
    mk(unif_d, 1);
    rep("uniform, no stores", run<LANES, false>(tab, idx, n_idx, out));
    rep("uniform, stores evict_first", run<LANES, true>(tab, idx, n_idx, out));
    mk(rank_d, 0);
    rep("zipf SCATTERED hot rows, no stores", run<LANES, false>(tab, idx, n_idx, out));
    rep("zipf SCATTERED hot rows, stores evict_first", run<LANES, true>(tab, idx, n_idx, out));
    mk(rank_d, 1);
    rep("zipf PACKED (popularity order), no stores", run<LANES, false>(tab, idx, n_idx, out));
    rep("zipf PACKED (popularity order), stores evict_first", run<LANES, true>(tab, idx, n_idx, out));
}

int main() {
    const long long rows = 4400000, n_idx = 29500000;
    const double alpha = 0.8;
    float4 *tab, *out; int *idx, *rank_d, *unif_d, *perm_d;
    CK(cudaMalloc(&tab, rows * 128)); CK(cudaMemset(tab, 0, rows * 128));
    CK(cudaMalloc(&out, (n_idx / 3 + 4096) * 128));
    CK(cudaMalloc(&idx, n_idx * 4)); CK(cudaMalloc(&rank_d, n_idx * 4)); CK(cudaMalloc(&unif_d, n_idx * 4));
    CK(cudaMalloc(&perm_d, rows * 4));
    std::vector<double> cdf(rows);
    double s = 0;
    for (long long r = 0; r < rows; ++r) { s += std::pow((double)(r + 1), -alpha); cdf[r] = s; }
    std::mt19937_64 rng(1234);
    std::uniform_real_distribution<double> uni(0.0, 1.0);
    std::vector<int> rank(n_idx), unif(n_idx), perm(rows);
    for (long long e = 0; e < n_idx; ++e) {
        rank[e] = (int)(std::upper_bound(cdf.begin(), cdf.end(), uni(rng) * s) - cdf.begin());
        unif[e] = (int)(rng() % (unsigned long long)rows);
    }
    for (long long r = 0; r < rows; ++r) perm[r] = (int)r;
    std::shuffle(perm.begin(), perm.end(), rng);
    CK(cudaMemcpy(perm_d, perm.data(), rows * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(rank_d, rank.data(), n_idx * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(unif_d, unif.data(), n_idx * 4, cudaMemcpyHostToDevice));
    sweep<4>(tab, idx, rank_d, unif_d, perm_d, n_idx, out);    // 64-byte rows (d/P = 16)
    sweep<8>(tab, idx, rank_d, unif_d, perm_d, n_idx, out);    // 128-byte rows (d/P = 32)
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
