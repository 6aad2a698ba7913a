// Microbenchmark: achievable bandwidth of random 512-byte row gathers (warp per row, float4 per
// lane) as a function of table size and loads in flight.  Design input for lgcn_spmm.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu && ./gather_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <int U>
__global__ void __launch_bounds__(256) gather_kernel(const float4* __restrict__ tab, const int* __restrict__ idx,
                                                     long long n_idx, float4* __restrict__ out, int rows_per_warp) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long base = warp * rows_per_warp;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int i = 0; i < rows_per_warp; i += U) {
        int r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] = (base + i + u < n_idx) ? __ldg(idx + base + i + u) : 0;
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = __ldg(tab + (size_t)r[u] * 32 + lane);
#pragma unroll
        for (int u = 0; u < U; ++u) { acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w; }
    }
    out[warp * 32 + lane] = acc;
}

template <int U>
float run(const float4* tab, const int* idx, long long n_idx, float4* out, int rpw) {
    long long warps = (n_idx + rpw - 1) / rpw;
    int blocks = (int)((warps * 32 + 255) / 256);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    gather_kernel<U><<<blocks, 256>>>(tab, idx, n_idx, out, rpw);
    cudaEventRecord(a);
    for (int it = 0; it < 3; ++it) gather_kernel<U><<<blocks, 256>>>(tab, idx, n_idx, out, rpw);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 3;
}

int main() {
    const long long n_idx = 32LL << 20;   // 32 M gathered rows = 16 GiB of traffic
    int* idx; cudaMalloc(&idx, n_idx * 4);
    float4* out; cudaMalloc(&out, (n_idx / 8 + 1024) * 512);
    std::vector<int> h(n_idx);
    for (double gb : {0.125, 0.5, 2.0, 7.5, 30.0}) {
        long long rows = (long long)(gb * 1e9 / 512);
        float4* tab; if (cudaMalloc(&tab, rows * 512) != cudaSuccess) { printf("alloc fail %.1f\n", gb); continue; }
        cudaMemset(tab, 0, rows * 512);
        unsigned long long s = 88172645463325252ULL;
        for (long long i = 0; i < n_idx; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % (unsigned long long)rows); }
        cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
        for (int rpw : {8, 64}) {
            float m1 = run<1>(tab, idx, n_idx, out, rpw), m4 = run<4>(tab, idx, n_idx, out, rpw), m8 = run<8>(tab, idx, n_idx, out, rpw);
            double by = (double)n_idx * 512 / 1e9;
            printf("table %.3f GB rows/warp %d : U=1 %.0f GB/s  U=4 %.0f GB/s  U=8 %.0f GB/s\n", gb, rpw, by / m1 * 1e3, by / m4 * 1e3, by / m8 * 1e3);
        }
        cudaFree(tab);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
