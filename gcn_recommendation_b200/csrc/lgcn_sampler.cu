// lgcn_sampler.cu -- device-side BPR batch sampler (sm_100a).
//
// Replaces BPRDataset + DataLoader(shuffle=True) of reference main.py:349-363, 462-464: python
// `random.randint` rejection sampling in 4 worker processes becomes the bottleneck once a training
// step costs < 1 ms.  Semantics kept: an epoch visits every training
// interaction exactly once in a random order (a keyed bijection of [0, E), so no permutation
// array is stored), and the negative is uniform over the items the user has not interacted with
// (rejection against the user's sorted CSR row).  Parity is statistical, not bitwise: the
// reference's stream depends on python's RNG and its worker count.
#include "lgcn_common.cuh"

namespace lgcn {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {      // splitmix64 finaliser
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// keyed bijection of [0, n): 4-round Feistel network on 2*hb bits with cycle walking
__device__ __forceinline__ uint64_t permute(uint64_t i, uint64_t n, int hb, uint64_t key) {
    const uint64_t mask = (1ull << hb) - 1ull;
    do {
        uint64_t l = i >> hb, r = i & mask;
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            const uint64_t f = mix64(r ^ (key + 0x632be59bd9b4e019ull * (uint64_t)(round + 1))) & mask;
            const uint64_t t = l ^ f;
            l = r;
            r = t;
        }
        i = (l << hb) | r;
    } while (i >= n);
    return i;
}

// state[0] = epoch, state[1] = position inside the epoch
__global__ void __launch_bounds__(256)
sample_bpr_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                  int64_t num_users, int64_t num_items, int64_t n_edges, uint64_t seed,
                  const int64_t *__restrict__ state, int64_t bs, int64_t *__restrict__ users,
                  int64_t *__restrict__ pos, int64_t *__restrict__ neg) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= bs) return;
    int64_t epoch = state[0], at = state[1] + s;
    while (at >= n_edges) { at -= n_edges; ++epoch; }          // batch may straddle the epoch end
    int hb = 1;
    while ((1ull << (2 * hb)) < (uint64_t)n_edges) ++hb;
    const uint64_t e = permute((uint64_t)at, (uint64_t)n_edges, hb, mix64(seed ^ mix64((uint64_t)epoch)));
    // the user-side rows [0, num_users) of the graph CSR are exactly the training interactions
    int64_t lo = 0, hi = num_users;                            // last row with rowptr[row] <= e
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if ((uint64_t)(uint32_t)__ldg(rowptr + mid) <= e) lo = mid; else hi = mid;
    }
    const int64_t u = lo;
    const int beg = __ldg(rowptr + u), end = __ldg(rowptr + u + 1);
    users[s] = u;
    pos[s] = (int64_t)__ldg(col + e) - num_users;
    uint64_t ctr = mix64(seed ^ 0x5851f42d4c957f2dull) ^ ((uint64_t)epoch << 40) ^ (uint64_t)at;
    for (int attempt = 0; attempt < 64; ++attempt) {
        ctr = mix64(ctr + (uint64_t)attempt);
        const int64_t cand = (int64_t)__umul64hi(ctr, (uint64_t)num_items);          // uniform in [0, I)
        const int key = (int)(cand + num_users);
        int a = beg, b = end;                                   // binary search in the sorted row
        while (a < b) {
            const int m = (a + b) >> 1;
            if (__ldg(col + m) < key) a = m + 1; else b = m;
        }
        if (!(a < end && __ldg(col + a) == key) || attempt == 63) { neg[s] = cand; break; }
    }
}

__global__ void sample_advance_kernel(int64_t *state, int64_t n_edges, int64_t bs) {
    int64_t at = state[1] + bs, epoch = state[0];
    while (at >= n_edges) { at -= n_edges; ++epoch; }
    state[0] = epoch;
    state[1] = at;
}

}  // namespace lgcn

extern "C" int lgcn_sample_bpr(const int32_t *rowptr, const int32_t *col, int64_t num_users,
                               int64_t num_items, uint64_t seed, int64_t *state, int64_t bs,
                               int64_t *users, int64_t *pos, int64_t *neg, int64_t n_edges,
                               lgcn_stream_t stream) {
    if (!rowptr || !col || !state || !users || !pos || !neg) return LGCN_E_BAD_ARG;
    if (num_users <= 0 || num_items <= 0 || bs <= 0 || n_edges <= 0) return LGCN_E_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    lgcn::sample_bpr_kernel<<<(unsigned)((bs + 255) / 256), 256, 0, st>>>(
        rowptr, col, num_users, num_items, n_edges, seed, state, bs, users, pos, neg);
    LGCN_LAUNCH_CHECK();
    lgcn::sample_advance_kernel<<<1, 1, 0, st>>>(state, n_edges, bs);
    LGCN_LAUNCH_CHECK();
    return 0;
}
