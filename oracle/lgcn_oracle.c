/*
 * lgcn_oracle.c -- CPU restatement of the reference's LightGCN hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gcn_recommendation_b200/,
 * models/) may link, load or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, and only as the checker.
 *
 * The reference (Validation-m3sSAGE/GCN_Recommendation) is pure Python; its arithmetic
 * lives in PyTorch 2.11 / SciPy 1.18 / NumPy 2.3 (unpinned upstream, these are the
 * versions the golden vectors under tests/golden/ were produced with).  Each function
 * below restates what those library calls compute at the cited reference call site.
 * Parity status: PINNED against outputs of the reference itself run in the build
 * container (the .npz files in tests/golden, written by oracle/make_golden.py).
 *
 * Plain C99, scalar, single thread.  Build: see oracle/Makefile
 * (-ffp-contract=off so only the explicit fmaf() calls fuse).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * Y = A_hat * X     reference: models/lightgcn.py:45, models/lightgcn_fusion.py:56
 * torch.sparse.mm on CPU is bit-equal to a per-row sequential fp32 FMA in ascending
 * column order (verified by tests/test_oracle_golden.py).
 * CSR: rowptr[N+1] int64, col[nnz] int32 (ascending inside a row), val[nnz] fp32.
 * ---------------------------------------------------------------------------------- */
void lgcn_oracle_spmm(const int64_t *rowptr, const int32_t *col, const float *val,
                      const float *X, float *Y, int64_t n_rows, int32_t d)
{
    for (int64_t r = 0; r < n_rows; ++r) {
        float *y = Y + r * (int64_t)d;
        for (int32_t j = 0; j < d; ++j) y[j] = 0.0f;
        for (int64_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
            const float w = val[e];
            const float *x = X + (int64_t)col[e] * d;
            for (int32_t j = 0; j < d; ++j) y[j] = fmaf(w, x[j], y[j]);
        }
    }
}

/* ------------------------------------------------------------------------------------
 * F = mean(stack(E_0..E_K))   reference: models/lightgcn.py:54
 * bit-equal to the sequential sum E_0+E_1+... followed by a true division by K+1.
 * layers: (K+1) contiguous blocks of n floats.
 * ---------------------------------------------------------------------------------- */
void lgcn_oracle_layer_mean(const float *layers, int32_t n_layers, int64_t n, float *out)
{
    const float div = (float)n_layers;
    for (int64_t i = 0; i < n; ++i) {
        float s = layers[i];
        for (int32_t l = 1; l < n_layers; ++l) s = s + layers[(int64_t)l * n + i];
        out[i] = s / div;
    }
}

/* ------------------------------------------------------------------------------------
 * BPR + L2 loss and its gradients.   reference: main.py:366-402 with the gathers of
 * main.py:496-497.   loss = -mean(log(sigmoid(pos-neg)+1e-8)) + lam*(|u0|^2+|p0|^2+|n0|^2)/B
 *
 * F   : [N,d] propagated table (users first, items at row item_offset+i)
 * E0u : [U,d] layer-0 user table, E0i : [I,d] layer-0 item (id) table -- the 4th/5th
 *       return values of forward (models/lightgcn.py:81, models/lightgcn_fusion.py:65)
 * gF  : [N,d]  += dLoss/dF        (caller zeroes)
 * gE0u: [U,d], gE0i: [I,d]  += dLoss/dE0 through the regulariser only (caller zeroes)
 * Accumulation in double; the result is compared at 1e-5 relative, not bit-exactly
 * (torch's vectorised reductions are not order-pinned).
 * Returns the scalar loss.
 * ---------------------------------------------------------------------------------- */
double lgcn_oracle_bpr(const float *F, const float *E0u, const float *E0i,
                       const int64_t *users, const int64_t *pos, const int64_t *neg,
                       int64_t bs, int32_t d, int64_t item_offset, float lam,
                       float *gF, float *gE0u, float *gE0i)
{
    double bpr = 0.0, reg = 0.0;
    const double invB = 1.0 / (double)bs;
    for (int64_t s = 0; s < bs; ++s) {
        const float *fu = F + users[s] * (int64_t)d;
        const float *fp = F + (item_offset + pos[s]) * (int64_t)d;
        const float *fn = F + (item_offset + neg[s]) * (int64_t)d;
        float ps = 0.0f, ns = 0.0f;
        for (int32_t j = 0; j < d; ++j) { ps += fu[j] * fp[j]; ns += fu[j] * fn[j]; }
        const float x = ps - ns;
        const float sg = 1.0f / (1.0f + expf(-x));
        bpr += -(double)logf(sg + 1e-8f);
        /* d/dx of -log(sigmoid(x)+1e-8) = -sg(1-sg)/(sg+1e-8) */
        const float coef = (float)(-(double)sg * (1.0 - (double)sg) / ((double)sg + 1e-8) * invB);
        if (gF) {
            float *gu = gF + users[s] * (int64_t)d;
            float *gp = gF + (item_offset + pos[s]) * (int64_t)d;
            float *gn = gF + (item_offset + neg[s]) * (int64_t)d;
            for (int32_t j = 0; j < d; ++j) {
                gu[j] += coef * (fp[j] - fn[j]);
                gp[j] += coef * fu[j];
                gn[j] -= coef * fu[j];
            }
        }
        const float *eu = E0u + users[s] * (int64_t)d;
        const float *ep = E0i + pos[s] * (int64_t)d;
        const float *en = E0i + neg[s] * (int64_t)d;
        const float c2 = (float)(2.0 * (double)lam * invB);
        for (int32_t j = 0; j < d; ++j) {
            reg += (double)eu[j] * eu[j] + (double)ep[j] * ep[j] + (double)en[j] * en[j];
        }
        if (gE0u) {
            float *hu = gE0u + users[s] * (int64_t)d;
            float *hp = gE0i + pos[s] * (int64_t)d;
            float *hn = gE0i + neg[s] * (int64_t)d;
            for (int32_t j = 0; j < d; ++j) {
                hu[j] += c2 * eu[j]; hp[j] += c2 * ep[j]; hn[j] += c2 * en[j];
            }
        }
    }
    return bpr * invB + (double)lam * reg * invB;
}

/* ------------------------------------------------------------------------------------
 * Dense Adam (no weight decay, no amsgrad).   reference: main.py:469,526
 * torch.optim.Adam single-tensor formula: bias corrections in double on the host,
 * step_size = lr/bc1, denom = sqrt(v)/sqrt(bc2) + eps, p -= step_size * m/denom.
 * t is the 1-based step count AFTER increment.
 * ---------------------------------------------------------------------------------- */
void lgcn_oracle_adam(float *p, const float *g, float *m, float *v, int64_t n, int64_t t,
                      float lr, float beta1, float beta2, float eps)
{
    const double bc1 = 1.0 - pow((double)beta1, (double)t);
    const double bc2 = 1.0 - pow((double)beta2, (double)t);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    for (int64_t i = 0; i < n; ++i) {
        const float gi = g[i];
        m[i] = m[i] + (gi - m[i]) * (1.0f - beta1);               /* lerp_ */
        v[i] = v[i] * beta2 + (1.0f - beta2) * gi * gi;           /* mul_ + addcmul_ */
        const float denom = sqrtf(v[i]) / bc2_sqrt + eps;
        p[i] = p[i] - step_size * (m[i] / denom);
    }
}

/* ------------------------------------------------------------------------------------
 * Fusion item block.   reference: models/lightgcn_fusion.py:45-49
 * H = leaky_relu([E_id | C] W^T + b, 0.01);  W is [d, d+c] row-major (nn.Linear).
 * Accumulates in double (checked at 1e-5, sgemm order is not pinned).
 * pre (optional) receives the pre-activation.
 * ---------------------------------------------------------------------------------- */
void lgcn_oracle_fusion_fwd(const float *Eid, const float *C, const float *W, const float *b,
                            int64_t n_items, int32_t d, int32_t c, float *H, float *pre)
{
    const int32_t kin = d + c;
    for (int64_t i = 0; i < n_items; ++i) {
        for (int32_t o = 0; o < d; ++o) {
            double acc = (double)b[o];
            const float *w = W + (int64_t)o * kin;
            for (int32_t k = 0; k < d; ++k) acc += (double)Eid[i * d + k] * w[k];
            for (int32_t k = 0; k < c; ++k) acc += (double)C[i * (int64_t)c + k] * w[d + k];
            const float h = (float)acc;
            if (pre) pre[i * d + o] = h;
            H[i * d + o] = h > 0.0f ? h : 0.01f * h;
        }
    }
}

/* backward of the fusion block: given gH (grad wrt the activated output) and the
 * pre-activation, produce gEid [I,d], gW [d,d+c], gb [d] (all overwritten). */
void lgcn_oracle_fusion_bwd(const float *Eid, const float *C, const float *W, const float *pre,
                            const float *gH, int64_t n_items, int32_t d, int32_t c,
                            float *gEid, float *gW, float *gb)
{
    const int32_t kin = d + c;
    double *aW = (double *)calloc((size_t)d * kin, sizeof(double));
    double *ab = (double *)calloc((size_t)d, sizeof(double));
    for (int64_t i = 0; i < n_items; ++i) {
        for (int32_t k = 0; k < d; ++k) gEid[i * d + k] = 0.0f;
        for (int32_t o = 0; o < d; ++o) {
            const float g = gH[i * d + o] * (pre[i * d + o] > 0.0f ? 1.0f : 0.01f);
            ab[o] += g;
            const float *w = W + (int64_t)o * kin;
            double *aw = aW + (int64_t)o * kin;
            for (int32_t k = 0; k < d; ++k) {
                gEid[i * d + k] += g * w[k];
                aw[k] += (double)g * Eid[i * d + k];
            }
            for (int32_t k = 0; k < c; ++k) aw[d + k] += (double)g * C[i * (int64_t)c + k];
        }
    }
    for (int64_t j = 0; j < (int64_t)d * kin; ++j) gW[j] = (float)aW[j];
    for (int32_t o = 0; o < d; ++o) gb[o] = (float)ab[o];
    free(aW); free(ab);
}

/* ------------------------------------------------------------------------------------
 * Full-rank rating: scores, train-item mask, top-k.   reference: main.py:420-426
 * score(u,i) = sum_j Fu[u,j]*Fi[i,j] as a sequential fp32 FMA over j (the order the
 * exact re-score of the CUDA path uses; MKL's sgemm order is not pinned, so ids are
 * compared bit-exactly against THIS order and against the reference wherever
 * neighbouring scores are not fp32 near-ties).  Masked items get -1e10 BEFORE the
 * top-k (main.py:422-424).  Order: score descending, ties -> lower item id first
 * (torch.topk's tie order is implementation-defined).
 * users: row indices into Fu.  mask CSR is indexed by position in `users`.
 * out_ids [nu,k] int32, out_scores [nu,k] fp32.
 * ---------------------------------------------------------------------------------- */
void lgcn_oracle_score_topk(const float *Fu, const float *Fi, const int64_t *users, int64_t nu,
                            int64_t n_items, int32_t d, const int64_t *mask_rowptr,
                            const int32_t *mask_col, int32_t k, int32_t *out_ids,
                            float *out_scores)
{
    float *s = (float *)malloc((size_t)n_items * sizeof(float));
    for (int64_t q = 0; q < nu; ++q) {
        const float *fu = Fu + users[q] * (int64_t)d;
        for (int64_t i = 0; i < n_items; ++i) {
            const float *fi = Fi + i * (int64_t)d;
            float acc = 0.0f;
            for (int32_t j = 0; j < d; ++j) acc = fmaf(fu[j], fi[j], acc);
            s[i] = acc;
        }
        if (mask_rowptr)
            for (int64_t e = mask_rowptr[q]; e < mask_rowptr[q + 1]; ++e) s[mask_col[e]] = -1e10f;
        /* k rounds of selection; k is 20 */
        for (int32_t r = 0; r < k; ++r) {
            int64_t best = -1; float bv = 0.0f;
            for (int64_t i = 0; i < n_items; ++i) {
                const float v = s[i];
                if (isnan(v)) continue;
                if (best < 0 || v > bv) { best = i; bv = v; }
            }
            if (best < 0) { out_ids[q * k + r] = -1; out_scores[q * k + r] = NAN; continue; }
            out_ids[q * k + r] = (int32_t)best;
            out_scores[q * k + r] = bv;
            s[best] = NAN;
        }
    }
    free(s);
}
