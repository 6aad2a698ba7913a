// lgcn_graph.cu -- COO -> CSR and symmetric-normalisation weights (sm_100a).
//
// Replaces the scipy COO->CSR conversion and D^-1/2 A D^-1/2 product of reference
// main.py:321-331, and the per-call coalesce/sort torch.sparse.mm performs on the
// uncoalesced COO tensor built at reference main.py:334-336.  Integer work: bit-exact.
#include "lgcn_common.cuh"

namespace lgcn {

// status[0] += number of idx[i] outside [lo, hi): the IndexError the reference's gathers raise
// (reference main.py:496-497 index the model outputs with the batch) as a device-side count.
__global__ void check_indices_kernel(const int64_t *__restrict__ idx, int64_t n, int64_t lo, int64_t hi,
                                     int32_t *__restrict__ status) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t v = idx[i];
        bad += (v < lo || v >= hi) ? 1 : 0;
    }
    if (bad) atomicAdd(status, bad);
}

// rowptr[r] = first e with coo_row[e] >= r.  One thread per entry boundary.
__global__ void csr_from_sorted_coo_kernel(const int64_t *__restrict__ row,
                                           const int64_t *__restrict__ colin, int64_t nnz,
                                           int64_t n_rows, int64_t n_cols, int32_t *__restrict__ rowptr,
                                           int32_t *__restrict__ col, int32_t *__restrict__ status) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e <= nnz; e += stride) {
        const int64_t r_prev = (e == 0) ? -1 : row[e - 1];
        const int64_t r_cur = (e == nnz) ? n_rows : row[e];
        if (e < nnz) {
            const int64_t c = colin[e];
            col[e] = (int32_t)c;
            bool bad = r_cur < 0 || r_cur >= n_rows || c < 0 || c >= n_cols;
            if (e > 0) bad = bad || r_cur < r_prev || (r_cur == r_prev && c <= colin[e - 1]);
            if (bad) atomicAdd(status, 1);
        }
        // rows r_prev+1 .. r_cur start at e (empty rows included)
        for (int64_t r = max(r_prev + 1, (int64_t)0); r <= min(r_cur, n_rows); ++r) rowptr[r] = (int32_t)e;
    }
}

__global__ void edge_weights_kernel(const int32_t *__restrict__ rowptr,
                                    const int32_t *__restrict__ col,
                                    const float *__restrict__ dinv, const float *__restrict__ mult,
                                    float *__restrict__ val, int64_t n_rows) {
    // one warp per row: lanes stride over the row's entries
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_rows) return;
    const int beg = rowptr[warp], end = rowptr[warp + 1];
    const float dr = dinv[warp];
    for (int e = beg + lane; e < end; e += 32) {
        const float m = mult ? mult[e] : 1.0f;
        val[e] = __fmul_rn(__fmul_rn(dr, m), dinv[col[e]]);
    }
}

}  // namespace lgcn

extern "C" int lgcn_check_indices(const int64_t *idx, int64_t n, int64_t lo, int64_t hi, int32_t *status,
                                  lgcn_stream_t stream) {
    if (n < 0 || !status || (n > 0 && !idx)) return LGCN_E_BAD_ARG;
    if (n == 0) return 0;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    lgcn::check_indices_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        idx, n, lo, hi, status);
    LGCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int lgcn_csr_from_sorted_coo(const int64_t *coo_row, const int64_t *coo_col,
                                        int64_t nnz, int64_t n_rows, int64_t n_cols, int32_t *rowptr,
                                        int32_t *col, int32_t *status, lgcn_stream_t stream) {
    if (nnz < 0 || n_rows < 0 || n_cols < 0 || !rowptr || !status) return LGCN_E_BAD_ARG;
    if (n_cols > 0x7fffffffLL) return LGCN_E_TOO_LARGE;
    if (nnz > 0 && (!coo_row || !coo_col || !col)) return LGCN_E_BAD_ARG;
    if (nnz > 0x7fffffffLL || n_rows > 0x7fffffffLL) return LGCN_E_TOO_LARGE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return (int)e;
    const int threads = 256;
    const int64_t blocks = (nnz + 1 + threads - 1) / threads;
    const unsigned grid = (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);
    lgcn::csr_from_sorted_coo_kernel<<<grid, threads, 0, st>>>(coo_row, coo_col, nnz, n_rows, n_cols,
                                                              rowptr, col, status);
    LGCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int lgcn_edge_weights(const int32_t *rowptr, const int32_t *col, const float *dinv,
                                 const float *mult, float *val, int64_t n_rows,
                                 lgcn_stream_t stream) {
    if (n_rows < 0 || !rowptr || !dinv) return LGCN_E_BAD_ARG;
    if (n_rows == 0) return 0;
    if (!col || !val) return LGCN_E_BAD_ARG;
    if (n_rows > 0x7fffffffLL / 32 * 8) return LGCN_E_TOO_LARGE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int threads = 256;
    const int64_t blocks = (n_rows * 32 + threads - 1) / threads;
    lgcn::edge_weights_kernel<<<(unsigned)blocks, threads, 0, st>>>(rowptr, col, dinv, mult, val,
                                                                   n_rows);
    LGCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int lgcn_abi_version(void) { return LGCN_ABI_VERSION; }

extern "C" const char *lgcn_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case LGCN_E_BAD_DIM: return "lgcn: embedding dim must be one of 16/32/64/128/256";
        case LGCN_E_BAD_ARG: return "lgcn: bad argument (null pointer, negative size or bad mode)";
        case LGCN_E_TOO_LARGE: return "lgcn: size exceeds an int32 index limit";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "lgcn: unknown error";
    }
}
