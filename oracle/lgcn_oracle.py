"""CPU oracle for the LightGCN hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker.  The product
path (``gcn_recommendation_b200/``, ``models/``) never does.

numpy restates the integer / index work (adjacency build, reference ``main.py:283-336``);
``lgcn_oracle.c`` (plain C, loaded with ctypes) restates the floating-point loops.  Every
function cites the reference lines it follows.  Parity status: PINNED against outputs of
the reference itself run in the build container -- ``tests/golden/*.npz`` written by
``oracle/make_golden.py`` and checked by ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblgcn_oracle.so")
_lib = None


def build(force=False):
    """Compile lgcn_oracle.c with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "lgcn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liblgcn_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.lgcn_oracle_bpr.restype = ctypes.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ----------------------------------------------------------------------------------------
# a1  adjacency build + symmetric normalisation        reference main.py:283-336
# ----------------------------------------------------------------------------------------
def build_norm_adj(train_user, train_item, num_users, num_items, num_brands,
                   item_brand=None):
    """Restates reference ``main.py:283-336``.

    rows = [u ; i+U], cols = [i+U ; u] of ones (``:304-313``); scipy's
    ``diags(d).dot(A).dot(diags(d))`` lands in CSR with duplicates summed and columns
    ascending; ``d = np.power(rowsum, -0.5)`` in fp32 with inf -> 0 (``:326-329``); the
    value of entry (r,c) with multiplicity m is ``fl32(fl32(d_r*m)*d_c)`` (``:330-331``).
    ``item_brand`` = optional (item_idx, brand_idx) arrays for the tripartite graph
    (``:300-306``).

    Returns dict(rowptr int64[N+1], col int32[nnz], val fp32[nnz], dinv fp32[N],
    deg fp32[N] (weighted row sums), mult fp32[nnz]).
    """
    U, I, B = int(num_users), int(num_items), int(num_brands)
    N = U + I + B
    u = _i64(train_user)
    it = _i64(train_item) + U
    rows = [u, it]
    cols = [it, u]
    if item_brand is not None:
        ib_i = _i64(item_brand[0]) + U
        ib_b = _i64(item_brand[1]) + U + I
        rows += [ib_i, ib_b]
        cols += [ib_b, ib_i]
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    key = rows * N + cols
    ukey, mult = np.unique(key, return_counts=True)          # sorted (row, col), multiplicity
    r = ukey // N
    c = (ukey % N).astype(np.int32)
    multf = mult.astype(np.float32)
    # rowsum of the COO of fp32 ones (main.py:326): sums of small integers are exact in fp32
    deg = np.bincount(r, weights=mult, minlength=N).astype(np.float32)
    with np.errstate(divide="ignore"):
        dinv = np.power(deg, np.float32(-0.5)).astype(np.float32)   # main.py:328 (fp32 powf)
    dinv[np.isinf(dinv)] = 0.0
    val = ((dinv[r] * multf).astype(np.float32) * dinv[c]).astype(np.float32)
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(np.bincount(r, minlength=N), out=rowptr[1:])
    return dict(rowptr=rowptr, col=c, val=val, dinv=dinv, deg=deg, mult=multf)


# ----------------------------------------------------------------------------------------
# a2  propagation + layer mean          reference models/lightgcn.py:37-59
# ----------------------------------------------------------------------------------------
def spmm(rowptr, col, val, X):
    X = _f32(X)
    N, d = X.shape
    Y = np.empty_like(X)
    lib().lgcn_oracle_spmm(_p(_i64(rowptr)), _p(_i32(col)), _p(_f32(val)), _p(X), _p(Y),
                           ctypes.c_int64(N), ctypes.c_int32(d))
    return Y


def propagate(rowptr, col, val, E0, n_layers):
    """Returns (F, [E_0..E_K]): K sequential-FMA SpMMs then mean(stack) as sequential sum
    and a true division by K+1 (reference ``models/lightgcn.py:44-54``)."""
    rowptr, col, val = _i64(rowptr), _i32(col), _f32(val)
    layers = [_f32(E0)]
    for _ in range(n_layers):
        layers.append(spmm(rowptr, col, val, layers[-1]))
    stack = np.ascontiguousarray(np.stack(layers, 0))
    F = np.empty_like(layers[0])
    lib().lgcn_oracle_layer_mean(_p(stack), ctypes.c_int32(n_layers + 1),
                                 ctypes.c_int64(layers[0].size), _p(F))
    return F, layers


# ----------------------------------------------------------------------------------------
# a3  backward of the propagation (autograd of models/lightgcn.py:44-54)
# ----------------------------------------------------------------------------------------
def propagate_backward(rowptr, col, val, gF, n_layers):
    """dL/dE0 = sum_k A^k g/(K+1), evaluated Horner-style acc <- g' + A acc (A symmetric)."""
    g1 = (_f32(gF) / np.float32(n_layers + 1)).astype(np.float32)
    acc = g1.copy()
    for _ in range(n_layers):
        acc = g1 + spmm(rowptr, col, val, acc)
    return acc


# ----------------------------------------------------------------------------------------
# a4  BPR + L2                         reference main.py:366-402, 496-497
# ----------------------------------------------------------------------------------------
def bpr_loss(F, E0_user, E0_item, users, pos, neg, num_users, lam, want_grads=True):
    """Returns (loss, gF, gE0_user, gE0_item); the g* are None if not want_grads."""
    F, E0_user, E0_item = _f32(F), _f32(E0_user), _f32(E0_item)
    users, pos, neg = _i64(users), _i64(pos), _i64(neg)
    d = F.shape[1]
    gF = np.zeros_like(F) if want_grads else None
    gU = np.zeros_like(E0_user) if want_grads else None
    gI = np.zeros_like(E0_item) if want_grads else None
    loss = lib().lgcn_oracle_bpr(_p(F), _p(E0_user), _p(E0_item), _p(users), _p(pos), _p(neg),
                                 ctypes.c_int64(len(users)), ctypes.c_int32(d),
                                 ctypes.c_int64(num_users), ctypes.c_float(lam),
                                 _p(gF), _p(gU), _p(gI))
    return float(loss), gF, gU, gI


def bpr_brand_term(F, users, pos, neg, item_to_brand, num_users, num_items):
    """The brand / author BPR term of reference ``main.py:382-391``:
    ``-mean(log(sigmoid(<F_u, F_brand[b(pos)]> - <F_u, F_brand[b(neg)]>) + 1e-8))`` with
    ``b = item_to_brand`` (the lookup ``main.py:499-511`` intends).  Same arithmetic as the item
    term with the brand block as the "item" table and no regulariser, so the same C loop serves:
    item offset = U + I, lambda = 0.  Returns (term, gF) -- unweighted; the caller applies
    ``brand_loss_weight`` (``main.py:401``)."""
    F = _f32(F)
    i2b = _i64(item_to_brand)
    bp, bn = i2b[_i64(pos)], i2b[_i64(neg)]
    off = int(num_users) + int(num_items)
    gF = np.zeros_like(F)
    dummy_u = np.zeros_like(F[:num_users])          # lambda = 0: regulariser rows are never used
    dummy_b = np.zeros_like(F[off:])
    term = lib().lgcn_oracle_bpr(_p(F), _p(dummy_u), _p(dummy_b), _p(_i64(users)), _p(bp), _p(bn),
                                 ctypes.c_int64(len(bp)), ctypes.c_int32(F.shape[1]),
                                 ctypes.c_int64(off), ctypes.c_float(0.0), _p(gF), None, None)
    return float(term), gF


# ----------------------------------------------------------------------------------------
# a5  Adam                              reference main.py:469,526
# ----------------------------------------------------------------------------------------
def adam_step(p, g, m, v, t, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """In-place on p, m, v (fp32 C-contiguous). t = 1-based step count."""
    for a in (p, m, v):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    g = _f32(g)
    lib().lgcn_oracle_adam(_p(p), _p(g), _p(m), _p(v), ctypes.c_int64(p.size),
                           ctypes.c_int64(t), ctypes.c_float(lr), ctypes.c_float(beta1),
                           ctypes.c_float(beta2), ctypes.c_float(eps))


# ----------------------------------------------------------------------------------------
# a6  fusion projection                 reference models/lightgcn_fusion.py:45-49
# ----------------------------------------------------------------------------------------
def fusion_forward(E_id, C, W, b):
    E_id, C, W, b = _f32(E_id), _f32(C), _f32(W), _f32(b)
    n, d = E_id.shape
    c = C.shape[1]
    H = np.empty((n, d), np.float32)
    pre = np.empty((n, d), np.float32)
    lib().lgcn_oracle_fusion_fwd(_p(E_id), _p(C), _p(W), _p(b), ctypes.c_int64(n),
                                 ctypes.c_int32(d), ctypes.c_int32(c), _p(H), _p(pre))
    return H, pre


def fusion_backward(E_id, C, W, pre, gH):
    E_id, C, W, pre, gH = _f32(E_id), _f32(C), _f32(W), _f32(pre), _f32(gH)
    n, d = E_id.shape
    c = C.shape[1]
    gE = np.empty((n, d), np.float32)
    gW = np.empty((d, d + c), np.float32)
    gb = np.empty((d,), np.float32)
    lib().lgcn_oracle_fusion_bwd(_p(E_id), _p(C), _p(W), _p(pre), _p(gH), ctypes.c_int64(n),
                                 ctypes.c_int32(d), ctypes.c_int32(c), _p(gE), _p(gW), _p(gb))
    return gE, gW, gb


# ----------------------------------------------------------------------------------------
# a7  full-rank rating + metrics         reference main.py:404-439
# ----------------------------------------------------------------------------------------
def mask_csr(users, train_user, train_item, num_users):
    """Per-evaluated-user sorted list of train items (``train_df.groupby(user).apply(list)``,
    reference ``main.py:407``), as a CSR indexed by position in ``users``."""
    tu, ti = _i64(train_user), _i64(train_item)
    order = np.lexsort((ti, tu))
    tu, ti = tu[order], ti[order]
    start = np.searchsorted(tu, np.arange(num_users + 1))
    users = _i64(users)
    cnt = start[users + 1] - start[users]
    rowptr = np.zeros(len(users) + 1, np.int64)
    np.cumsum(cnt, out=rowptr[1:])
    idx = np.repeat(start[users] - rowptr[:-1], cnt) + np.arange(rowptr[-1])
    return rowptr, ti[idx].astype(np.int32)


def score_topk(F_user, F_item, users, mask_rowptr, mask_col, k=20):
    F_user, F_item = _f32(F_user), _f32(F_item)
    users = _i64(users)
    ids = np.empty((len(users), k), np.int32)
    sc = np.empty((len(users), k), np.float32)
    lib().lgcn_oracle_score_topk(_p(F_user), _p(F_item), _p(users), ctypes.c_int64(len(users)),
                                 ctypes.c_int64(F_item.shape[0]), ctypes.c_int32(F_user.shape[1]),
                                 _p(None if mask_rowptr is None else _i64(mask_rowptr)),
                                 _p(None if mask_col is None else _i32(mask_col)),
                                 ctypes.c_int32(k), _p(ids), _p(sc))
    return ids, sc


def recall_ndcg(topk_ids, targets):
    """hit = target in top-k; recall = mean(hit); ndcg = mean(hit / log2(pos+2)) in float64
    (reference ``main.py:430-439``)."""
    topk_ids = np.asarray(topk_ids)
    targets = np.asarray(targets).reshape(-1, 1)
    eq = topk_ids == targets
    hit = eq.any(1)
    posn = eq.argmax(1)
    ndcg = np.where(hit, 1.0 / np.log2(posn + 2.0), 0.0)
    return float(np.mean(hit.astype(np.float64))), float(np.mean(ndcg))


def eval_pairs(val_user, val_item):
    """``dict(zip(user, item))`` semantics of reference ``main.py:406-408``: the last row wins
    for a duplicated user, users keep first-appearance order."""
    d = dict(zip(np.asarray(val_user).tolist(), np.asarray(val_item).tolist()))
    return np.fromiter(d.keys(), np.int64, len(d)), np.fromiter(d.values(), np.int64, len(d))
