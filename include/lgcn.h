/*
 * lgcn.h -- C ABI of the B200-native LightGCN hot path (liblgcn_b200.so).
 *
 * The reference (Validation-m3sSAGE/GCN_Recommendation) is pure Python and has no FFI of
 * its own: its hot path is a chain of PyTorch / SciPy library calls.  Every entry point
 * below replaces one of those call sites (cited as reference file:line) and is what a
 * maintainer binds with ctypes from the reference's model plugin.
 *
 * Conventions
 *  - plain C types only; every pointer is a BORROWED DEVICE pointer unless it is named
 *    *_host; the library never allocates, frees, synchronises or throws.
 *  - every call takes the cudaStream_t to launch on (pass torch's current stream).
 *  - return value: 0 on success, a positive cudaError_t if a launch failed, or a negative
 *    LGCN_E_* code for an argument the kernels do not support.
 *  - fp32 everywhere; row-major tables [rows, d]; d in {16,32,64,128,256}; 16-byte aligned.
 *  - a "table" is the concatenation users | items | brands (reference models/lightgcn.py:40).
 */
#ifndef LGCN_H
#define LGCN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *lgcn_stream_t; /* == cudaStream_t */

#if defined(__GNUC__)
#define LGCN_API __attribute__((visibility("default")))
#else
#define LGCN_API
#endif

#define LGCN_ABI_VERSION 7

#define LGCN_E_BAD_DIM   (-1) /* embedding dim not supported              */
#define LGCN_E_BAD_ARG   (-2) /* null pointer / negative size / bad mode  */
#define LGCN_E_TOO_LARGE (-3) /* size exceeds an int32 index limit        */

LGCN_API int lgcn_abi_version(void);
/* human readable text for a return code of this library (static storage) */
LGCN_API const char *lgcn_error_string(int code);

/* ---------------------------------------------------------------------------------------
 * a1  Graph: COO -> CSR and symmetric normalisation.
 * Replaces: scipy coo->csr + D^-1/2 A D^-1/2 (reference main.py:321-331) and the COO tensor
 * torch.sparse.mm re-sorts on every call (reference main.py:334-336, models/lightgcn.py:45).
 * ------------------------------------------------------------------------------------- */

/* Row-major sorted COO (what main.py:334-336 hands to forward) -> CSR.
 * rowptr[n_rows+1], col[nnz] int32.  status[0] receives the number of entries that violate
 * strict (row,col) ordering or lie outside [0,n_rows) x [0,n_cols) (0 == valid CSR order, no
 * duplicates, every index in range). */
LGCN_API int lgcn_csr_from_sorted_coo(const int64_t *coo_row, const int64_t *coo_col, int64_t nnz,
                             int64_t n_rows, int64_t n_cols, int32_t *rowptr, int32_t *col,
                             int32_t *status, lgcn_stream_t stream);

/* status[0] += number of idx[i] outside [lo, hi) -- the device-side form of the IndexError the
 * reference's gathers raise for a bad batch / evaluation index (reference main.py:496-497,420).
 * The caller zeroes status and reads it back when it wants the verdict. */
LGCN_API int lgcn_check_indices(const int64_t *idx, int64_t n, int64_t lo, int64_t hi, int32_t *status,
                       lgcn_stream_t stream);

/* val[e] = fl32(fl32(dinv[row]*mult[e]) * dinv[col[e]])  (mult == NULL -> 1), the value
 * scipy produces at main.py:330-331.  dinv is computed by the caller with the same
 * np.power(fp32, -0.5) call as main.py:328 so that the weights are bit-equal. */
LGCN_API int lgcn_edge_weights(const int32_t *rowptr, const int32_t *col, const float *dinv,
                      const float *mult, float *val, int64_t n_rows, lgcn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * a2/a3/a5  Normalised-adjacency SpMM with fused epilogues.
 * Replaces: torch.sparse.mm (reference models/lightgcn.py:45, lightgcn_fusion.py:56), its
 * autograd backward, torch.mean(torch.stack()) (lightgcn.py:54) and optimizer.step()
 * (main.py:526) when fused into the last backward hop.
 *
 * y[r,:] = sum_e val[e] * X[col[e],:]  for local rows r in [0,n_rows); col indexes rows of X
 * (the full table), every epilogue array is indexed by the LOCAL row.
 *
 * Graph layout (built once per graph by the host, see graph.py: NormAdjCSR.plan):
 *  - rows with at most `long_row_threshold` entries ("short rows") live in `colval`, an array
 *    of packed {int32 col, float val} pairs in CSR order; rowptr[r] & 0x7fffffff is the first
 *    entry of row r.  A short row is accumulated as a SEQUENTIAL fp32 FMA in ascending column
 *    order -- bit-equal to the CPU reference (torch.sparse.mm).
 *  - bit 31 of rowptr[r] marks row r as a long row: it has no entries in `colval`; its
 *    entries live in `long_colval` (CSR over the long rows, `long_rowptr`), are cut into
 *    segments of seg_len entries, one sub-warp per segment, partial sums staged in seg_ws and
 *    combined in segment order (deterministic, not bit-equal to the sequential order).
 *  - L2 residency classes (optional, set by the graph plan when the tables exceed L2; honoured
 *    under LGCN_SPMM_F_STREAM_HINTS, always masked off): bit 31 of colval[e].col marks a HOT
 *    column -- one of the highest-degree nodes, as many as the L2 budget holds -- whose gathered
 *    row is gathered with evict_last (a preference only: measured +0.5..2 % at the Amazon shape;
 *    an L2 persisting set-aside on top of it made the streams slower); bit 30 marks a column referenced exactly once per launch
 *    (degree 1), whose row leaves L2 first.  Column ids therefore need n_cols <= 2^30.
 *    `long_colval` carries no class bits.
 * ------------------------------------------------------------------------------------- */
typedef struct lgcn_colval {
    int32_t col;
    float   val;
} lgcn_colval;
#define LGCN_COL_HOT  ((int32_t)0x80000000)
#define LGCN_COL_ONCE 0x40000000
#define LGCN_COL_MASK 0x3fffffff

typedef struct lgcn_spmm_args {
    const uint32_t    *rowptr;  /* [n_rows+1] entry offsets, bit 31 = long-row flag       */
    const lgcn_colval *colval;  /* entries of the short rows                              */
    const float   *x;           /* [*, d] gathered table                                  */
    int64_t        n_rows;
    int32_t        d;
    int32_t        mode;        /* LGCN_SPMM_*                                            */
    float         *y;           /* [n_rows,d] output (modes PLAIN, ADD, MEAN)             */
    const float   *addend;      /* ADD, ADAM: y = addend + A x ; may be NULL in ADAM      */
    const float   *layers[8];   /* MEAN: earlier layers E_0..E_{n-1} (local rows)         */
    int32_t        n_layers;    /* MEAN: number of entries in layers                      */
    int32_t        n_long;      /* number of long rows (0 = none)                         */
    const int32_t *long_row_ids;/* [n_long] local row ids, ascending                      */
    const int32_t *long_rowptr; /* [n_long+1] offsets into long_colval                    */
    const lgcn_colval *long_colval;
    const int32_t *long_seg_ptr;/* [n_long+1] first segment of each long row              */
    int32_t        seg_len;     /* entries per segment                                    */
    int32_t        n_seg;       /* total segments == long_seg_ptr[n_long] (host copy)     */
    float         *seg_ws;      /* [n_seg, d] scratch                                     */
    /* ADAM: g = addend + A x (+ addend2) is consumed by a dense Adam update of p           */
    const float   *addend2;     /* optional second addend (regulariser gradient)          */
    float         *p, *m, *v;   /* [n_rows,d] parameters and Adam moments                 */
    const float   *adam_scalars;/* device [2]: step_size=lr/(1-b1^t), sqrt(1-b2^t)        */
    float          beta1, beta2, eps;
    float         *g_out;       /* ADAM: optional [n_rows,d] copy of g (NULL = none)      */
    int32_t        flags;       /* LGCN_SPMM_F_*                                          */
    /* Sparse-gradient shortcuts of the backward pass (all optional, NULL = dense):
     * x_rowflag[c] == 0 means row c of x is all zero, so its gather is skipped (mode ADD only;
     * the first Horner hop reads g', which has at most 3*batch non-zero rows);
     * addend_rowflag[r] == 0 means row r of addend (and addend2) is all zero and is not read;
     * this array must be 4-byte aligned and padded to n_rows rounded up to a multiple of 16.
     * zero_row: a zero-filled device buffer of >= d floats the skipped loads are redirected to
     * (required when either flag array is given). */
    const uint8_t *x_rowflag;
    const uint8_t *addend_rowflag;
    const float   *zero_row;
    /* Sparse OUTPUT (mode ADD with x_rowflag and addend_rowflag only, optional): the kernel sets
     * y_rowflag[r] = 1 for the rows that summed a flagged row of x or have a flagged addend
     * (and for long rows), 0 for the others, and does NOT write the all-zero rows of y -- the
     * consumer must read y under y_rowflag (it is the next hop's x_rowflag).  [n_rows]. */
    uint8_t       *y_rowflag;
    /* Layer-0 override (LightGCN_Fusion: the item rows of layer 0 are the projected rows H, the
     * other rows the raw tables -- reference models/lightgcn_fusion.py:45-52 builds that table with
     * torch.cat every forward).  Rows c in [alt_begin, alt_begin + alt_rows) are read from
     * x_alt[(c - alt_begin) * d ...] instead of the main table:
     *   LGCN_SPMM_F_ALT_X       for the gathered rows of x (dense gathers only, no x_rowflag; not ADAM),
     *   LGCN_SPMM_F_ALT_LAYER0  for layers[0] of the MEAN epilogue.
     * NULL / no flag = none. */
    const float   *x_alt;
    int64_t        alt_begin, alt_rows;
    /* ADAM only: rows r in [skip_begin, skip_begin + skip_rows) are NOT updated (their layer-0 rows
     * are produced by the fusion projection, not by p); their gradient  addend + A x  (addend2 is
     * not added) is stored to g_skip[(r - skip_begin) * d ...] for the projection's backward. */
    float         *g_skip;
    int64_t        skip_begin, skip_rows;
    /* Work plan, both optional (NULL).
     * chunk_order: a permutation of the ceil(n_rows / R) chunks of R consecutive rows, R =
     *   lgcn_spmm_chunk_rows(n_rows, d, flags); worker w takes chunk chunk_order[w].  Tables narrower
     *   than 128 floats put several workers in one warp, which then walks to the longest of their
     *   chunks: an order that puts chunks of similar entry count next to each other (sorted inside
     *   windows, so that the streams stay local) removes that idling.  Results do not depend on it.
     *   Ignored when lgcn_spmm_chunk_rows() returns 0.
     * long_done (used on small, L2-resident graphs only): n_long zero-initialised counters.  When
     *   given, the worker that stores the LAST segment partial of a long row combines the row's
     *   partials (in segment order, as the combine launch would) and runs its epilogue, then
     *   re-arms the counter: no separate combine launch.  One concurrent lgcn_spmm call per
     *   counter array. */
    const int32_t *chunk_order;
    int32_t       *long_done;
} lgcn_spmm_args;

/* tables do not fit L2: stream entries / outputs / epilogue operands with L2 evict_first so
 * that they do not displace the gathered rows (set by the host when n_cols*d*4 >> L2) */
#define LGCN_SPMM_F_STREAM_HINTS 1
/* kernel selection overrides for tests and A/B measurements (default: chosen by d and mode) */
#define LGCN_SPMM_F_NO_RING 2       /* register-batch chunk kernel only                      */
#define LGCN_SPMM_F_BIG_PATH 4      /* large-graph kernels even when the graph is small      */
#define LGCN_SPMM_F_COLD_FIRST 8    /* gathers of unclassified columns use evict_first too   */
#define LGCN_SPMM_F_FORCE_RING 16   /* (no-op since ABI v5: the ring kernel serves ADAM too)  */
#define LGCN_SPMM_F_NO_PREFETCH 32  /* no L2 prefetch of the epilogue operands (A/B only)    */
#define LGCN_SPMM_F_ALT_X 64        /* gathers of rows in the alt range read x_alt           */
#define LGCN_SPMM_F_ALT_LAYER0 128  /* MEAN: layers[0] rows in the alt range read x_alt      */
#define LGCN_SPMM_F_LONG_DONE 256   /* lgcn_spmm_launches() only: the call will pass long_done */

#define LGCN_SPMM_PLAIN 0 /* y = A x                                                    */
#define LGCN_SPMM_ADD   1 /* y = addend + A x              (Horner backward hop)        */
#define LGCN_SPMM_MEAN  2 /* y = (E_0 + ... + E_{n-1} + A x) / (n+1), sequential sum then
                             a true division (reference models/lightgcn.py:54)          */
#define LGCN_SPMM_ADAM  3 /* g = addend + A x + addend2 ; Adam(p,m,v,g)                 */

LGCN_API int lgcn_spmm(const lgcn_spmm_args *args_host, lgcn_stream_t stream);
/* sizeof(lgcn_spmm_args) as compiled, so that a binding can verify its struct layout */
LGCN_API size_t lgcn_sizeof_spmm_args(void);
/* Host-only query: rows per chunk that lgcn_spmm_args.chunk_order permutes for this graph size,
 * width and flags; 0 when the kernel that will run ignores chunk_order; negative LGCN_E_* code. */
LGCN_API int lgcn_spmm_chunk_rows(int64_t n_rows, int32_t d, int32_t flags);
/* Host-only query (launches nothing): how many kernels one lgcn_spmm call launches for a graph of
 * n_rows rows with n_long long rows at width d under `flags` (1 = main kernel; with long rows
 * 3 = segments + main + combine, on small graphs 2 -- the segment workers ride in the main launch --
 * or 1 with LGCN_SPMM_F_LONG_DONE).  *small_path (optional) = 1 when the small-graph 4-row-chunk
 * path is taken.  Returns the count (> 0) or a negative LGCN_E_* code. */
LGCN_API int lgcn_spmm_launches(int64_t n_rows, int32_t d, int32_t n_long, int32_t flags,
                       int32_t *small_path);
/* Host-only query: the name of the MAIN kernel lgcn_spmm selects for this shape / mode (what a
 * profiler shows; bench.py labels its roofline with it).  sparse_x != 0: the call passes
 * x_rowflag.  Writes a NUL-terminated string of at most buf_bytes bytes; returns 0 or LGCN_E_*. */
LGCN_API int lgcn_spmm_kernel_name(int64_t n_rows, int32_t d, int32_t mode, int32_t flags,
                          int32_t sparse_x, char *buf, size_t buf_bytes);

/* ---------------------------------------------------------------------------------------
 * a4  Fused BPR + L2 step.
 * Replaces: the six row gathers (reference main.py:496-497), bpr_loss_reg (main.py:366-402)
 * and their autograd backward (index_put accumulate).
 *   x_s   = <F[u_s], F[io+p_s]> - <F[u_s], F[io+n_s]>
 *   loss  = -mean_s log(sigmoid(x_s)+1e-8) + lam*sum_s(|P[u_s]|^2+|P[io+p_s]|^2+|P[io+n_s]|^2)/bs
 *   gF   += grad_scale * dloss/dF            (rows u, io+p, io+n; float atomics)
 *   gP   += dloss/dP through the regulariser (2*lam/bs * row, once per occurrence)
 *           (+ the gF contribution too when LGCN_BPR_GP_INCLUDES_GF is set)
 * F: [N,d] propagated table; P: [N,d] layer-0 id table (the 4th/5th forward outputs);
 * io = item_offset = num_users.  sample_ws: [2*bs] floats of scratch.  loss_out: [1].
 * rowflag (optional, [N] bytes): set to 1 for every row that received a gradient, so that the
 * backward SpMM hops can skip the all-zero rows of gF / gP (see lgcn_spmm_args.x_rowflag).
 * ------------------------------------------------------------------------------------- */
#define LGCN_BPR_GP_INCLUDES_GF 1
#define LGCN_BPR_NO_GRAD        2 /* loss only */

LGCN_API int lgcn_bpr_fused(const float *F, const float *P, const int64_t *users, const int64_t *pos,
                   const int64_t *neg, int64_t bs, int32_t d, int64_t item_offset, float lam,
                   float grad_scale, int32_t flags, float *sample_ws, float *loss_out,
                   float *gF, float *gP, uint8_t *rowflag, lgcn_stream_t stream);

/* Feature-sharded tables (every rank owns d/P columns of every row): the step is split around
 * one small all-reduce.  lgcn_bpr_partial writes this rank's partial sums dots[0:bs]=<u,p>,
 * dots[bs:2bs]=<u,n>, dots[2bs:3bs]=|u0|^2+|p0|^2+|n0|^2; after summing dots over the ranks,
 * lgcn_bpr_apply forms the loss and scatters this rank's gradient columns (d = local width). */
LGCN_API int lgcn_bpr_partial(const float *F, const float *P, const int64_t *users, const int64_t *pos,
                     const int64_t *neg, int64_t bs, int32_t d, int64_t item_offset, float *dots,
                     lgcn_stream_t stream);
LGCN_API int lgcn_bpr_apply(const float *F, const float *P, const int64_t *users, const int64_t *pos,
                   const int64_t *neg, int64_t bs, int32_t d, int64_t item_offset, float lam,
                   float grad_scale, int32_t flags, const float *dots, float *sample_ws,
                   float *loss_out, float *gF, float *gP, uint8_t *rowflag, lgcn_stream_t stream);

/* zero the rows {u_s, io+p_s, io+n_s} of up to two [N,d] tables (undo of the scatter) */
LGCN_API int lgcn_zero_rows(float *t0, float *t1, uint8_t *rowflag, const int64_t *users,
                   const int64_t *pos, const int64_t *neg, int64_t bs, int32_t d,
                   int64_t item_offset, lgcn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * f1  Device-side BPR batch sampler.
 * Replaces: BPRDataset.__getitem__ + DataLoader(shuffle=True) (reference main.py:349-363,
 * 462-464).  rowptr/col: the UNFLAGGED CSR of the graph (rows [0,num_users) hold the training
 * interactions, columns = num_users + item, ascending); n_edges = rowptr[num_users].
 * state (device int64[2]) = {epoch, position in epoch}; every call emits the next bs
 * interactions of a keyed random permutation of the epoch (each interaction exactly once per
 * epoch) with a uniform negative the user has not interacted with, then advances state.
 * ------------------------------------------------------------------------------------- */
LGCN_API int lgcn_sample_bpr(const int32_t *rowptr, const int32_t *col, int64_t num_users,
                    int64_t num_items, uint64_t seed, int64_t *state, int64_t bs, int64_t *users,
                    int64_t *pos, int64_t *neg, int64_t n_edges, lgcn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * a5  Dense Adam (torch.optim.Adam defaults, reference main.py:469,526).
 * lgcn_adam_tick: t += 1 on the device and refresh adam_scalars = {lr/(1-b1^t),
 * sqrt(1-b2^t)} (step counter and scalars stay on the device so a step is graph-capturable).
 * lgcn_adam: g = g0 (+ g1); m,v,p updated in place.
 * ------------------------------------------------------------------------------------- */
LGCN_API int lgcn_adam_tick(int64_t *step_dev, float *adam_scalars, float lr, float beta1, float beta2,
                   lgcn_stream_t stream);
LGCN_API int lgcn_adam(float *p, const float *g0, const float *g1, float *m, float *v, int64_t n,
              const float *adam_scalars, float beta1, float beta2, float eps,
              lgcn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * a6  Fusion item block (reference models/lightgcn_fusion.py:45-49).
 *   H = leaky_relu([E_id | C] W^T + b, 0.01)      W: [d, d+c] row-major (nn.Linear.weight)
 * The concatenation is never materialised.  Backward: gH is the gradient w.r.t. H;
 * gEid [n_items,d] is overwritten; gW [d,d+c] and gb [d] are ACCUMULATED (caller zeroes).
 * ------------------------------------------------------------------------------------- */
LGCN_API int lgcn_fusion_proj_fwd(const float *Eid, const float *C, const float *W, const float *b,
                         int64_t n_items, int32_t d, int32_t c, float *H,
                         lgcn_stream_t stream);
/* test / debugging hook: 1 = force the fp32 SIMT kernels instead of the tcgen05 3xTF32 path */
LGCN_API void lgcn_fusion_force_simt(int on);
LGCN_API int lgcn_fusion_proj_bwd(const float *Eid, const float *C, const float *W, const float *H,
                         const float *gH, int64_t n_items, int32_t d, int32_t c, float *gEid,
                         float *gW, float *gb, lgcn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * a7  Full-rank rating: scores + train-item mask + top-k, scores never reach HBM.
 * Replaces: torch.matmul (reference main.py:420), the per-user mask loop (main.py:422-424)
 * and torch.topk (main.py:426).
 *   score(q,i) = <Fu[users[q]], Fi[i]>  (fp32, sequential FMA over the feature index)
 *   items in mask row q (CSR over query position, ascending item ids) are excluded
 *   out_ids/out_scores [nu,k]: descending score, ties -> lower item id first.
 * k <= 32.  workspace: lgcn_score_topk_workspace() bytes.
 * lgcn_eval_metrics: sums[0] += #hits, sums[1] += sum 1/log2(rank+2) (main.py:430-438).
 * ------------------------------------------------------------------------------------- */
LGCN_API size_t lgcn_score_topk_workspace(int64_t nu, int64_t n_items, int32_t d, int32_t k);
LGCN_API int lgcn_score_topk(const float *Fu, const float *Fi, const int64_t *users, int64_t nu,
                    int64_t n_items, int32_t d, const int64_t *mask_rowptr,
                    const int32_t *mask_col, int32_t k, int32_t *out_ids, float *out_scores,
                    void *workspace, size_t workspace_bytes, lgcn_stream_t stream);
/* Tensor-core path of the same operation for large catalogues (d = 64 or 128): a bf16
 * tcgen05.mma filter keeps 96 candidates per user (train items skipped in the epilogue), an
 * exact fp32 re-score orders the top k, and fail[q] = 1 marks users whose result is not
 * CERTIFIED exact (k-th exact score within the bf16 error bound of the filter threshold);
 * the caller re-runs those through lgcn_score_topk.  lgcn_score_tc_prepare converts the item
 * table once per table (bf16, UMMA canonical tiles) into the head of the workspace. */
/* A run with fewer user tiles than SMs (the reference rates 1024 users per batch, main.py:415)
 * cuts the catalogue into item splits so that the chip is filled; every split keeps its own
 * candidates, the ordered per-split lists are merged exactly and certified against the largest
 * threshold of any split.  lgcn_score_tc_workspace(nu, ..) is enough for EVERY call with at most
 * nu users (a sweep reuses one workspace and one prepared table for all of its user batches).
 * lgcn_score_tc_launches: host-only, kernels one lgcn_score_tc_topk call launches (2 or 3). */
LGCN_API size_t lgcn_score_tc_workspace(int64_t nu, int64_t n_items, int32_t d);
LGCN_API int lgcn_score_tc_launches(int64_t nu, int64_t n_items);
LGCN_API int lgcn_score_tc_prepare(const float *Fi, int64_t n_items, int32_t d, void *workspace,
                          size_t workspace_bytes, lgcn_stream_t stream);
LGCN_API int lgcn_score_tc_topk(const float *Fu, const float *Fi, const int64_t *users, int64_t nu,
                       int64_t n_items, int32_t d, const int64_t *mask_rowptr,
                       const int32_t *mask_col, int32_t k, int32_t *out_ids, float *out_scores,
                       int32_t *fail, void *workspace, size_t workspace_bytes,
                       lgcn_stream_t stream);
LGCN_API int lgcn_eval_metrics(const int32_t *topk_ids, const int64_t *targets, int64_t nu, int32_t k,
                      double *sums, lgcn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LGCN_H */
