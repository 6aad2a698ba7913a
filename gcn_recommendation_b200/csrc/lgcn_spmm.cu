// lgcn_spmm.cu -- C entry points of the normalised-adjacency SpMM (include/lgcn.h: lgcn_spmm,
// lgcn_spmm_kernel_name, lgcn_spmm_launches): argument validation and the dispatch on the table
// width.  The kernels live in lgcn_spmm_impl.cuh and are compiled per width in
// lgcn_spmm_d{16,32,64,128,256}.cu so that the five widths build in parallel.
#include "lgcn_spmm_impl.cuh"

#define LGCN_SPMM_DECLARE_WIDTH(D)                                                             \
    extern "C" __attribute__((visibility("hidden"))) int lgcn_spmm_launch_d##D(                \
        const lgcn_spmm_args *a, cudaStream_t st);
LGCN_SPMM_DECLARE_WIDTH(16)
LGCN_SPMM_DECLARE_WIDTH(32)
LGCN_SPMM_DECLARE_WIDTH(64)
LGCN_SPMM_DECLARE_WIDTH(128)
LGCN_SPMM_DECLARE_WIDTH(256)

extern "C" int lgcn_spmm(const lgcn_spmm_args *args, lgcn_stream_t stream) {
    using namespace lgcn;
    if (!args) return LGCN_E_BAD_ARG;
    const lgcn_spmm_args &a = *args;
    if (!dim_supported(a.d)) return LGCN_E_BAD_DIM;
    if (a.n_rows < 0 || !a.rowptr || !a.x) return LGCN_E_BAD_ARG;
    if (a.n_rows > 0x7fffffffLL) return LGCN_E_TOO_LARGE;
    if (a.n_rows > 0 && !a.colval) return LGCN_E_BAD_ARG;
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
    if (a.n_long < 0) return LGCN_E_BAD_ARG;
    if ((a.x_rowflag || a.addend_rowflag) && !a.zero_row) return LGCN_E_BAD_ARG;
    if (a.x_rowflag && a.mode != LGCN_SPMM_ADD) return LGCN_E_BAD_ARG;
    if (a.y_rowflag && !(a.x_rowflag && a.addend_rowflag)) return LGCN_E_BAD_ARG;
    if (a.flags & (LGCN_SPMM_F_ALT_X | LGCN_SPMM_F_ALT_LAYER0)) {
        if (!a.x_alt || a.alt_begin < 0 || a.alt_rows < 0 || a.alt_begin + a.alt_rows > 0x7fffffffLL)
            return LGCN_E_BAD_ARG;
        if ((a.flags & LGCN_SPMM_F_ALT_X) && (a.x_rowflag || a.mode == LGCN_SPMM_ADAM)) return LGCN_E_BAD_ARG;
        if ((a.flags & LGCN_SPMM_F_ALT_LAYER0) && a.mode != LGCN_SPMM_MEAN) return LGCN_E_BAD_ARG;
    }
    if (a.g_skip && (a.mode != LGCN_SPMM_ADAM || a.skip_begin < 0 || a.skip_rows < 0 ||
                     a.skip_begin + a.skip_rows > a.n_rows))
        return LGCN_E_BAD_ARG;
    if (a.n_long > 0 && (!a.long_row_ids || !a.long_rowptr || !a.long_colval || !a.long_seg_ptr ||
                         !a.seg_ws || a.seg_len <= 0 || a.n_seg <= 0))
        return LGCN_E_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (a.d) {
        case 16:  return lgcn_spmm_launch_d16(&a, st);
        case 32:  return lgcn_spmm_launch_d32(&a, st);
        case 64:  return lgcn_spmm_launch_d64(&a, st);
        case 128: return lgcn_spmm_launch_d128(&a, st);
        case 256: return lgcn_spmm_launch_d256(&a, st);
    }
    return LGCN_E_BAD_DIM;
}

extern "C" size_t lgcn_sizeof_spmm_args(void) { return sizeof(lgcn_spmm_args); }

static bool small_for_dim(int64_t n_rows, int32_t d, int32_t flags) {
    using namespace lgcn;
    switch (d) {
        case 16:  return small_graph<16>(n_rows, flags);
        case 32:  return small_graph<32>(n_rows, flags);
        case 64:  return small_graph<64>(n_rows, flags);
        case 128: return small_graph<128>(n_rows, flags);
        case 256: return small_graph<256>(n_rows, flags);
    }
    return false;
}

extern "C" int lgcn_spmm_kernel_name(int64_t n_rows, int32_t d, int32_t mode, int32_t flags,
                                     int32_t sparse_x, char *buf, size_t buf_bytes) {
    using namespace lgcn;
    if (!dim_supported(d)) return LGCN_E_BAD_DIM;
    if (n_rows < 0 || !buf || buf_bytes == 0 || mode < LGCN_SPMM_PLAIN || mode > LGCN_SPMM_ADAM)
        return LGCN_E_BAD_ARG;
    static const char *const kMode[] = {"plain", "add", "mean", "adam"};
    const bool small = small_for_dim(n_rows, d, flags);
    const bool hint = (flags & LGCN_SPMM_F_STREAM_HINTS) != 0;
    const bool xf = mode == LGCN_SPMM_ADD && sparse_x != 0;
    const bool ring = ring_path(small, mode, flags);
    if (ring && xf) snprintf(buf, buf_bytes, "spmm_live_kernel<%d,%s>", d, hint ? "hints" : "nohints");
    else if (ring) snprintf(buf, buf_bytes, "spmm_ring_kernel<%d,%s,%s>", d, kMode[mode], hint ? "hints" : "nohints");
    else snprintf(buf, buf_bytes, "spmm_chunk_kernel<%d,%s,%s,%s%s>", d, kMode[mode],
                  small ? "4-row chunks" : "16-row chunks", (hint && !small) ? "hints" : "nohints",
                  xf ? ",sparse-x" : "");
    return 0;
}

extern "C" int lgcn_spmm_chunk_rows(int64_t n_rows, int32_t d, int32_t flags) {
    using namespace lgcn;
    if (n_rows < 0) return LGCN_E_BAD_ARG;
    switch (d) {
        case 16:  return order_chunk_rows<16>(n_rows, flags);
        case 32:  return order_chunk_rows<32>(n_rows, flags);
        case 64:  return order_chunk_rows<64>(n_rows, flags);
        case 128: return order_chunk_rows<128>(n_rows, flags);
        case 256: return order_chunk_rows<256>(n_rows, flags);
    }
    return LGCN_E_BAD_DIM;
}

extern "C" int lgcn_spmm_launches(int64_t n_rows, int32_t d, int32_t n_long, int32_t flags,
                                  int32_t *small_path) {
    using namespace lgcn;
    if (!dim_supported(d)) return LGCN_E_BAD_DIM;
    if (n_rows < 0 || n_long < 0) return LGCN_E_BAD_ARG;
    const bool small = small_for_dim(n_rows, d, flags);
    if (small_path) *small_path = small ? 1 : 0;
    if (n_long == 0) return 1;                 // main kernel only
    if (small) return (flags & LGCN_SPMM_F_LONG_DONE) ? 1 : 2;   // segments ride in the main launch (+ combine)
    return 3;                                  // segments + main + combine
}
