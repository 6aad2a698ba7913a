"""Host-side operators over the C ABI (include/lgcn.h): one Python call per kernel entry point
and the ``torch.autograd.Function``s the drop-in models use.

Everything here launches hand-written sm_100a kernels; there is no eager/PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import functools
import os

import torch

from . import _lib
from ._lib import SPMM_ADAM, SPMM_ADD, SPMM_MEAN, SPMM_PLAIN, SpmmArgs, check, ptr, stream_ptr


# kernels of this library launched so far (bench.py's gpu_launches) and an optional per-launch
# CUDA-event profile of the SpMM (bench.py's roofline): set PROFILE to a list to collect
# (tag, start_event, end_event) on the launching stream.
COUNTERS = {"launches": 0}
CHUNK_ORDER_LARGE = os.environ.get("LGCN_CHUNK_ORDER_LARGE", "auto")   # auto | all | off, see _set_chunk_order
L2_STREAM_BYTES = 96 << 20      # tables larger than this are streamed with L2 evict_first hints
PROFILE = None
# L2 budget for the gathered rows of the highest-degree columns (gathered with evict_last; the rows
# of degree-1 columns leave first): LGCN_HOT_MB=0 turns the column classes off.  Measured at the
# Amazon shape (profiles/r01_hot_columns_sweep.txt): 32 MB is best, and worth only 0.5-2 % -- the
# L2 does not retain the hot rows under 6 TB/s of streaming (sector hit rate 11.5 -> 12.9 %).
HOT_BYTES = int(float(os.environ.get("LGCN_HOT_MB", "32")) * (1 << 20))
SPMM_FLAGS_EXTRA = int(os.environ.get("LGCN_SPMM_FLAGS", "0"))           # OR-ed into lgcn_spmm_args.flags (tests / A-B measurements force a kernel)


def _device_of(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a.device
        elif hasattr(a, "rowptr_flagged"):            # graph.NormAdjCSR
            return a.device
    return None


def on_device(fn):
    """Run an operator with the CUDA device of its first tensor / graph argument current: the
    library launches on the current context (cudaFuncSetAttribute, <<<>>>), so a tensor on cuda:1
    while cuda:0 is current must switch first.  One integer compare when they already agree."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = _device_of(args, kwargs)
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


class _timed:
    """``with _timed(tag):`` -- CUDA events around a launch when bench.py collects a PROFILE."""

    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if PROFILE is not None:
            self.s, self.e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.s.record()

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e.record()
            PROFILE.append((self.tag, self.s, self.e))
        return False


def _spmm_plan(n_rows, d, n_long, long_done=False):
    """(kernel launches per ``lgcn_spmm`` call, small-graph path?) as the library itself decides
    (host-only query, csrc/lgcn_spmm_impl.cuh ``launch_mode``).  ``long_done``: the calls carry
    the graph's long-row counters (small graphs then combine long rows inside the main launch)."""
    small = ctypes.c_int32(0)
    flags = int(SPMM_FLAGS_EXTRA) | (_lib.SPMM_F_LONG_DONE if long_done else 0)
    n = _lib.load().lgcn_spmm_launches(int(n_rows), int(d), int(n_long), flags, ctypes.byref(small))
    if n <= 0:
        check(n)
    return n, bool(small.value)


def _small_graph(n_rows, d):
    return _spmm_plan(n_rows, d, 0)[1]


def spmm_launches(g, d):
    """Kernels one ``lgcn_spmm`` call launches for this graph / width."""
    return _spmm_plan(g.n_rows, d, g.n_long, getattr(g, "long_done", None) is not None)[0]


def _launch_spmm(a, g, dev, tag):
    _set_chunk_order(a, g)
    n_kernels = spmm_launches(g, a.d)
    COUNTERS["launches"] += n_kernels
    if PROFILE is None:
        check(_lib.load().lgcn_spmm(ctypes.byref(a), stream_ptr(dev)))
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    check(_lib.load().lgcn_spmm(ctypes.byref(a), stream_ptr(dev)))
    e.record()
    PROFILE.append((tag, s, e))


def _spmm_args(g, x, mode, d):
    a = SpmmArgs()
    a.rowptr, a.colval = ptr(g.rowptr_flagged, "i32"), g.colval.data_ptr()
    a.x = ptr(x)
    a.n_rows, a.d, a.mode = g.n_rows, d, mode
    a.flags = SPMM_FLAGS_EXTRA
    if g.n_cols * d * 4 > L2_STREAM_BYTES:
        a.flags |= _lib.SPMM_F_STREAM_HINTS
    if a.flags & _lib.SPMM_F_STREAM_HINTS:
        n_hot = HOT_BYTES // (4 * d)
        if getattr(g, "n_hot", 0) != n_hot and (g.n_rows == g.n_cols or hasattr(g, "col_degree")):
            g.mark_hot_columns(n_hot)
    if g.n_long > 0:
        a.n_long = g.n_long
        a.long_row_ids = ptr(g.long_row_ids, "i32")
        a.long_rowptr = ptr(g.long_rowptr, "i32")
        a.long_colval = g.long_colval.data_ptr()
        a.long_seg_ptr = ptr(g.long_seg_ptr, "i32")
        a.seg_len, a.n_seg = g.seg_len, g.n_seg
        a.seg_ws = ptr(g.seg_ws(d))
        a.long_done = ptr(getattr(g, "long_done", None), "i32", allow_none=True)
    return a


def _set_chunk_order(a, g):
    """``lgcn_spmm_args.chunk_order`` (include/lgcn.h).  Small (L2-resident) graphs: always (Gowalla
    step -29 % together with the in-launch long-row combine).  HBM-streaming graphs with several
    workers per warp (d/P <= 64): the windowed order balances the workers but scatters each warp's
    chunks, which costs the coalescing of the entry stream -- measured per call at the Amazon shape
    (profiles/r02_chunk_order_ab.txt): launches with many streamed bytes per chunk (MEAN -3..-12 %,
    ADAM 0..-7 %) and the flagged walk with dense output (hop 2: -1..-15 %) gain at every width,
    pure gather launches (PLAIN / dense ADD at d/P = 16: +5..9 %) and the sparse-output hop (+10..18 %
    at d/P >= 32) lose.  ``auto`` follows that table; LGCN_CHUNK_ORDER_LARGE = all | off overrides."""
    if not hasattr(g, "chunk_order_for"):
        return
    small = _small_graph(g.n_rows, a.d)
    if not small:
        if CHUNK_ORDER_LARGE == "off":
            return
        if CHUNK_ORDER_LARGE != "all" and not (
                a.mode in (SPMM_MEAN, SPMM_ADAM) or (a.x_rowflag and not a.y_rowflag)):
            return
    rows = _lib.load().lgcn_spmm_chunk_rows(int(g.n_rows), int(a.d), int(a.flags))
    if rows > 0:
        a.chunk_order = ptr(g.chunk_order_for(rows, not small), "i32", allow_none=True)


def spmm_kernel_name(g, d, mode, sparse_x=False):
    """Name of the main kernel ``lgcn_spmm`` picks for this graph / width / mode (host-only query
    of the library's own selection, csrc/lgcn_spmm_impl.cuh ``launch_mode``; bench.py labels the
    roofline with it)."""
    code = {"plain": SPMM_PLAIN, "add": SPMM_ADD, "add_xs": SPMM_ADD, "add_xf": SPMM_ADD,
            "mean": SPMM_MEAN, "adam": SPMM_ADAM}[mode]
    flags = SPMM_FLAGS_EXTRA | (_lib.SPMM_F_STREAM_HINTS if g.n_cols * d * 4 > L2_STREAM_BYTES else 0)
    buf = ctypes.create_string_buffer(128)
    check(_lib.load().lgcn_spmm_kernel_name(int(g.n_rows), int(d), code, flags,
                                            int(sparse_x or mode in ("add_xs", "add_xf")), buf, 128))
    return buf.value.decode()


def _check_table(t, rows, d, name):
    if t.dim() != 2 or t.shape[0] < rows or t.shape[1] != d:
        raise _lib.LgcnError(f"{name}: expected at least [{rows},{d}], got {tuple(t.shape)}")


def _set_alt(a, alt, d, gather, layer0):
    """Layer-0 override: ``alt = (rows [n, d], first_row)`` -- those rows of the gathered table
    (``gather``) and / or of ``layers[0]`` (``layer0``) are read from ``rows`` instead."""
    rows, begin = alt
    if rows.dim() != 2 or rows.shape[1] != d:
        raise _lib.LgcnError("x_alt: expected [rows, d]")
    a.x_alt = ptr(rows)
    a.alt_begin, a.alt_rows = int(begin), int(rows.shape[0])
    a.flags |= (_lib.SPMM_F_ALT_X if gather else 0) | (_lib.SPMM_F_ALT_LAYER0 if layer0 else 0)


@on_device
def spmm(g, x, out=None, addend=None, mean_layers=None, x_rowflag=None, addend_rowflag=None,
         zero_row=None, y_rowflag=None, x_alt=None, layer0_alt=None):
    """out = A_hat x  (+ addend)  |  mean over [*mean_layers, A_hat x] (reference
    ``models/lightgcn.py:45,54``).  ``g``: :class:`graph.NormAdjCSR`; ``x`` [n_cols, d].

    Sparse-gradient shortcuts of the backward hops (include/lgcn.h): ``x_rowflag`` /
    ``addend_rowflag`` mark the non-zero rows of ``x`` / ``addend``; with ``y_rowflag`` the kernel
    reports which output rows can be non-zero and leaves the others UNWRITTEN (pass it as the
    next hop's ``x_rowflag``).

    ``x_alt = (rows, first_row)``: the LightGCN_Fusion layer 0 -- rows ``[first_row, first_row +
    len(rows))`` of ``x`` are gathered from ``rows`` instead (the projected item block), so the
    concatenated layer-0 table of reference ``models/lightgcn_fusion.py:52`` is never built;
    ``layer0_alt``: the same override for ``mean_layers[0]``."""
    d = x.shape[1]
    _check_table(x, g.n_cols, d, "x")
    if out is None:
        out = torch.empty((g.n_rows, d), dtype=torch.float32, device=x.device)
    _check_table(out, g.n_rows, d, "out")
    if out.data_ptr() == x.data_ptr():
        raise _lib.LgcnError("spmm cannot run in place")
    if mean_layers is not None:
        a = _spmm_args(g, x, SPMM_MEAN, d)
        if not 1 <= len(mean_layers) <= 8:
            raise _lib.LgcnError("mean epilogue supports 1..8 earlier layers")
        for i, l in enumerate(mean_layers):
            _check_table(l, g.n_rows, d, "layer")
            a.layers[i] = ptr(l)
        a.n_layers = len(mean_layers)
    elif addend is not None:
        _check_table(addend, g.n_rows, d, "addend")
        a = _spmm_args(g, x, SPMM_ADD, d)
        a.addend = ptr(addend)
        if x_rowflag is not None or addend_rowflag is not None:
            a.x_rowflag = ptr(x_rowflag, "u8", allow_none=True)
            a.addend_rowflag = ptr(addend_rowflag, "u8", allow_none=True)
            a.zero_row = ptr(zero_row)
            if y_rowflag is not None:
                if y_rowflag.numel() < g.n_rows:
                    raise _lib.LgcnError("y_rowflag: expected at least n_rows bytes")
                a.y_rowflag = ptr(y_rowflag, "u8")
    else:
        a = _spmm_args(g, x, SPMM_PLAIN, d)
    if x_alt is not None or layer0_alt is not None:
        if x_alt is not None and layer0_alt is not None and x_alt[0].data_ptr() != layer0_alt[0].data_ptr():
            raise _lib.LgcnError("x_alt and layer0_alt must be the same rows")
        if x_rowflag is not None:
            raise _lib.LgcnError("x_alt cannot be combined with x_rowflag")
        _set_alt(a, x_alt if x_alt is not None else layer0_alt, d, x_alt is not None,
                 layer0_alt is not None and mean_layers is not None)
    a.y = ptr(out)
    tag = ("plain", "add", "mean")[a.mode]
    if x_rowflag is not None:
        tag = "add_xs" if y_rowflag is not None else "add_xf"     # sparse in (+ sparse out)
    _launch_spmm(a, g, x.device, tag)
    return out


@on_device
def spmm_adam(g, x, p, m, v, adam_scalars, addend=None, addend2=None, betas=(0.9, 0.999),
              eps=1e-8, g_out=None, addend_rowflag=None, zero_row=None, skip=None):
    """Last backward hop fused with Adam: grad = addend + A_hat x + addend2; Adam(p, m, v, grad)
    (reference ``main.py:525-526``).  ``skip = (g_rows [n, d], first_row)``: those rows are not
    parameters of this table (LightGCN_Fusion item rows: produced by the projection) -- they are
    left alone and their gradient ``addend + A_hat x`` is written to ``g_rows``."""
    d = x.shape[1]
    a = _spmm_args(g, x, SPMM_ADAM, d)
    for t, n in ((p, "p"), (m, "m"), (v, "v")):
        _check_table(t, g.n_rows, d, n)
    a.p, a.m, a.v = ptr(p), ptr(m), ptr(v)
    a.addend = ptr(addend, allow_none=True)
    a.addend2 = ptr(addend2, allow_none=True)
    a.adam_scalars = ptr(adam_scalars)
    a.beta1, a.beta2, a.eps = betas[0], betas[1], eps
    a.g_out = ptr(g_out, allow_none=True)
    if addend_rowflag is not None:
        a.addend_rowflag = ptr(addend_rowflag, "u8")
        a.zero_row = ptr(zero_row)
    if skip is not None:
        rows, begin = skip
        if rows.dim() != 2 or rows.shape[1] != d or begin < 0 or begin + rows.shape[0] > g.n_rows:
            raise _lib.LgcnError("skip: expected ([rows, d], first_row) inside the table")
        a.g_skip = ptr(rows)
        a.skip_begin, a.skip_rows = int(begin), int(rows.shape[0])
    _launch_spmm(a, g, x.device, "adam")


def propagate(g, e0, n_layers, out=None, work=None, alt=None):
    """K-layer propagation with the layer mean fused into the last SpMM (reference
    ``models/lightgcn.py:44-54``).  Returns F [N,d].  ``work``: optional list of K-1 scratch
    tables (E_1..E_{K-1}).  ``alt = (rows, first_row)``: layer 0 is ``e0`` with those rows
    replaced (LightGCN_Fusion's projected item block) without materialising it."""
    if n_layers < 1:
        raise _lib.LgcnError("n_layers must be >= 1")
    layers = [e0]
    for k in range(n_layers - 1):
        buf = work[k] if work is not None else None
        layers.append(spmm(g, layers[-1], out=buf, x_alt=alt if k == 0 else None))
    return spmm(g, layers[-1], out=out, mean_layers=layers, x_alt=alt if n_layers == 1 else None,
                layer0_alt=alt)


def propagate_backward(g, grad_f, n_layers, work=None):
    """dL/dE0 = sum_k A^k g/(K+1) as Horner hops acc <- g' + A acc (A symmetric), the
    autograd of reference ``models/lightgcn.py:44-54``."""
    g1 = grad_f * (1.0 / (n_layers + 1))
    acc = g1
    for k in range(n_layers):
        buf = work[k % 2] if work is not None else None
        acc = spmm(g, acc, out=buf, addend=g1)
    return acc


@on_device
def bpr_fused(F, P, users, pos, neg, num_users, lam, grad_scale=1.0, gF=None, gP=None,
              gp_includes_gf=False, sample_ws=None, loss_out=None, rowflag=None):
    """Fused gather + BPR + L2 + scatter-add (reference ``main.py:366-402,496-497``)."""
    bs = users.numel()
    d = F.shape[1]
    dev = F.device
    if sample_ws is None:
        sample_ws = torch.empty(2 * bs, dtype=torch.float32, device=dev)
    if loss_out is None:
        loss_out = torch.empty(1, dtype=torch.float32, device=dev)
    flags = 0
    if gF is None and gP is None:
        flags |= _lib.BPR_NO_GRAD
    if gp_includes_gf:
        flags |= _lib.BPR_GP_INCLUDES_GF
    COUNTERS["launches"] += 2
    check(_lib.load().lgcn_bpr_fused(ptr(F), ptr(P), ptr(users, "i64"), ptr(pos, "i64"),
                                     ptr(neg, "i64"), bs, d, num_users, lam, grad_scale, flags,
                                     ptr(sample_ws), ptr(loss_out), ptr(gF, allow_none=True),
                                     ptr(gP, allow_none=True), ptr(rowflag, "u8", allow_none=True),
                                     stream_ptr(dev)))
    return loss_out


@on_device
def bpr_partial(F, P, users, pos, neg, num_users, dots):
    """Feature-sharded step, phase 1: this rank's partial <u,p>, <u,n>, |.|^2 per sample."""
    COUNTERS["launches"] += 1
    check(_lib.load().lgcn_bpr_partial(ptr(F), ptr(P), ptr(users, "i64"), ptr(pos, "i64"),
                                       ptr(neg, "i64"), users.numel(), F.shape[1], num_users,
                                       ptr(dots), stream_ptr(F.device)))
    return dots


@on_device
def bpr_apply(F, P, users, pos, neg, num_users, lam, dots, grad_scale=1.0, gF=None, gP=None,
              gp_includes_gf=False, sample_ws=None, loss_out=None, rowflag=None):
    """Feature-sharded step, phase 2: loss from the rank-summed dots, scatter local columns."""
    bs = users.numel()
    dev = F.device
    if sample_ws is None:
        sample_ws = torch.empty(2 * bs, dtype=torch.float32, device=dev)
    if loss_out is None:
        loss_out = torch.empty(1, dtype=torch.float32, device=dev)
    flags = 0
    if gF is None and gP is None:
        flags |= _lib.BPR_NO_GRAD
    if gp_includes_gf:
        flags |= _lib.BPR_GP_INCLUDES_GF
    COUNTERS["launches"] += 2
    check(_lib.load().lgcn_bpr_apply(ptr(F), ptr(P), ptr(users, "i64"), ptr(pos, "i64"),
                                     ptr(neg, "i64"), bs, F.shape[1], num_users, lam, grad_scale,
                                     flags, ptr(dots), ptr(sample_ws), ptr(loss_out),
                                     ptr(gF, allow_none=True), ptr(gP, allow_none=True),
                                     ptr(rowflag, "u8", allow_none=True), stream_ptr(dev)))
    return loss_out


@on_device
def zero_rows(t0, t1, users, pos, neg, num_users, rowflag=None):
    d = t0.shape[1]
    COUNTERS["launches"] += 1
    check(_lib.load().lgcn_zero_rows(ptr(t0), ptr(t1, allow_none=True),
                                     ptr(rowflag, "u8", allow_none=True), ptr(users, "i64"),
                                     ptr(pos, "i64"), ptr(neg, "i64"), users.numel(), d, num_users,
                                     stream_ptr(t0.device)))


@on_device
def sample_bpr(g, num_users, num_items, n_edges, seed, state, users, pos, neg):
    """Next batch of (user, pos, neg) triplets on the device (reference ``main.py:349-363``):
    one epoch = a keyed random permutation of the training interactions, negatives uniform over
    the user's non-interacted items.  ``state``: device int64[2] = {epoch, position}."""
    COUNTERS["launches"] += 2
    check(_lib.load().lgcn_sample_bpr(ptr(g.rowptr, "i32"), ptr(g.col, "i32"), num_users, num_items,
                                      seed, ptr(state, "i64"), users.numel(), ptr(users, "i64"),
                                      ptr(pos, "i64"), ptr(neg, "i64"), n_edges,
                                      stream_ptr(users.device)))


@on_device
def check_indices(status, *ranges):
    """status[0] += how many indices fall outside their range; ``ranges`` = (idx int64 tensor, lo,
    hi) triples.  The reference's gathers raise IndexError for those (main.py:496-497); here the
    caller reads ``status`` when it wants the verdict (no sync on the step path)."""
    lib = _lib.load()
    for idx, lo, hi in ranges:
        COUNTERS["launches"] += 1
        check(lib.lgcn_check_indices(ptr(idx, "i64"), idx.numel(), int(lo), int(hi),
                                     ptr(status, "i32"), stream_ptr(idx.device)))
    return status


@on_device
def adam_tick(step_dev, scalars, lr, betas=(0.9, 0.999)):
    COUNTERS["launches"] += 1
    check(_lib.load().lgcn_adam_tick(ptr(step_dev, "i64"), ptr(scalars), lr, betas[0], betas[1],
                                     stream_ptr(scalars.device)))


@on_device
def adam(p, g0, m, v, scalars, g1=None, betas=(0.9, 0.999), eps=1e-8):
    COUNTERS["launches"] += 1
    with _timed("adam_standalone"):
        check(_lib.load().lgcn_adam(ptr(p), ptr(g0), ptr(g1, allow_none=True), ptr(m), ptr(v),
                                    p.numel(), ptr(scalars), betas[0], betas[1], eps,
                                    stream_ptr(p.device)))


@on_device
def fusion_proj_fwd(e_id, content, W, b, out=None):
    """leaky_relu([E_id | C] W^T + b) without the concat (reference
    ``models/lightgcn_fusion.py:45-49``)."""
    n, d = e_id.shape
    c = content.shape[1]
    if W.shape != (d, d + c) or b.shape != (d,) or content.shape[0] != n:
        raise _lib.LgcnError("fusion_proj: shape mismatch")
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=e_id.device)
    COUNTERS["launches"] += 1
    with _timed("fusion_fwd"):
        check(_lib.load().lgcn_fusion_proj_fwd(ptr(e_id), ptr(content), ptr(W), ptr(b), n, d, c,
                                               ptr(out), stream_ptr(e_id.device)))
    return out


@on_device
def fusion_proj_bwd(e_id, content, W, H, gH, g_eid=None, gW=None, gb=None):
    n, d = e_id.shape
    c = content.shape[1]
    dev = e_id.device
    if g_eid is None:
        g_eid = torch.empty((n, d), dtype=torch.float32, device=dev)
    if gW is None:
        gW = torch.zeros((d, d + c), dtype=torch.float32, device=dev)
    if gb is None:
        gb = torch.zeros((d,), dtype=torch.float32, device=dev)
    COUNTERS["launches"] += 2
    with _timed("fusion_bwd"):
        check(_lib.load().lgcn_fusion_proj_bwd(ptr(e_id), ptr(content), ptr(W), ptr(H), ptr(gH), n, d,
                                               c, ptr(g_eid), ptr(gW), ptr(gb), stream_ptr(dev)))
    return g_eid, gW, gb


TC_MIN_ITEMS = 8192       # below this the exact SIMT kernel is used directly
STATS = {"tc_users": 0, "tc_fallback_users": 0}


@on_device
def score_topk_exact(F_user, F_item, users, mask_rowptr=None, mask_col=None, k=20):
    """Exact fp32 SIMT kernel (sequential FMA scores): the reference's arithmetic."""
    nu = users.numel()
    d = F_user.shape[1]
    dev = F_user.device
    ids = torch.empty((nu, k), dtype=torch.int32, device=dev)
    sc = torch.empty((nu, k), dtype=torch.float32, device=dev)
    if nu == 0:
        return ids, sc
    lib = _lib.load()
    wsb = lib.lgcn_score_topk_workspace(nu, F_item.shape[0], d, k)
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
    COUNTERS["launches"] += 2 if wsb else 1
    check(lib.lgcn_score_topk(ptr(F_user), ptr(F_item), ptr(users, "i64"), nu,
                              F_item.shape[0], d, ptr(mask_rowptr, "i64", allow_none=True),
                              ptr(mask_col, "i32", allow_none=True), k, ptr(ids, "i32"),
                              ptr(sc), ws.data_ptr(), wsb, stream_ptr(dev)))
    return ids, sc


def _sub_csr(rowptr, col, idx):
    """Rows ``idx`` of a CSR (int64 rowptr) as a compact CSR (host-free torch ops)."""
    cnt = rowptr[idx + 1] - rowptr[idx]
    rp = torch.zeros(idx.numel() + 1, dtype=torch.int64, device=rowptr.device)
    torch.cumsum(cnt, 0, out=rp[1:])
    src = torch.repeat_interleave(rowptr[idx] - rp[:-1], cnt) + torch.arange(int(rp[-1].item()),
                                                                            device=rowptr.device)
    return rp, col[src].to(torch.int32).contiguous()


class TcRater:
    """Prepared tensor-core rating of ONE item table (reference ``main.py:413-426`` rates every
    user batch against the same propagated table): the bf16 UMMA tiles + item norms are built once
    (``prepare``), and the workspace -- sized for ``max_users`` per call -- is allocated once and
    reused by every user batch of a sweep."""

    def __init__(self, n_items, d, device, max_users):
        lib = _lib.load()
        self.n_items, self.d, self.max_users, self.dev = int(n_items), int(d), int(max_users), device
        self.wsb = lib.lgcn_score_tc_workspace(self.max_users, self.n_items, self.d)
        if self.wsb == 0:
            raise _lib.LgcnError("tensor-core rating supports d = 64 or 128")
        self.ws = torch.empty(self.wsb, dtype=torch.uint8, device=device)
        self.prepared_for = None                      # (data_ptr, version) of the prepared table

    @on_device
    def prepare(self, F_item):
        if F_item.shape != (self.n_items, self.d):
            raise _lib.LgcnError("TcRater.prepare: item table shape changed")
        COUNTERS["launches"] += 2
        check(_lib.load().lgcn_score_tc_prepare(ptr(F_item), self.n_items, self.d, self.ws.data_ptr(),
                                                self.wsb, stream_ptr(self.dev)))
        self.prepared_for = (F_item.data_ptr(), F_item._version)

    @on_device
    def topk(self, F_user, F_item, users, mask_rowptr, mask_col, k, out_ids, out_sc, fail):
        """Filter + exact refine of one user batch (<= max_users) into the given output slices."""
        nu = users.numel()
        if nu > self.max_users:
            raise _lib.LgcnError("TcRater.topk: batch larger than the workspace was sized for")
        lib = _lib.load()
        COUNTERS["launches"] += lib.lgcn_score_tc_launches(nu, self.n_items)
        check(lib.lgcn_score_tc_topk(ptr(F_user), ptr(F_item), ptr(users, "i64"), nu, self.n_items,
                                     self.d, ptr(mask_rowptr, "i64", allow_none=True),
                                     ptr(mask_col, "i32", allow_none=True), k, ptr(out_ids, "i32"),
                                     ptr(out_sc), ptr(fail, "i32"), self.ws.data_ptr(), self.wsb,
                                     stream_ptr(self.dev)))


# users per tensor-core launch of a sweep: whole waves of 148 CTAs x 128 users (the CTAs of a wave
# stream the same item tiles, so the table is read from DRAM once per wave)
TC_WAVE_USERS = 148 * 128
TC_BATCH_WAVES = int(os.environ.get("LGCN_TC_BATCH_WAVES", "4"))


@on_device
def score_topk(F_user, F_item, users, mask_rowptr=None, mask_col=None, k=20, tensor_cores=None,
               rater=None, batch_users=None):
    """Full-rank scores + train mask + top-k (reference ``main.py:420-426``).
    Returns (ids int32 [nu,k], scores fp32 [nu,k]); ids are exactly those of fp32 sequential-FMA
    scoring.  Large catalogues go through the tcgen05 filter + exact re-score in user batches of
    ``batch_users`` (one prepared table and one workspace for the whole sweep: pass a
    :class:`TcRater` to keep them across calls); users whose result is not certified exact are
    re-run by the exact kernel at the end (one host sync per sweep)."""
    nu = users.numel()
    d = F_user.shape[1]
    n_items = F_item.shape[0]
    dev = F_user.device
    if tensor_cores is None:
        tensor_cores = d in (64, 128) and n_items >= TC_MIN_ITEMS and k <= 32
    if not tensor_cores or nu == 0:
        return score_topk_exact(F_user, F_item, users, mask_rowptr, mask_col, k)
    if batch_users is None:
        batch_users = TC_WAVE_USERS * TC_BATCH_WAVES
    batch_users = max(1, min(int(batch_users), nu))
    if rater is None:
        rater = TcRater(n_items, d, dev, batch_users)
    elif (rater.n_items, rater.d) != (n_items, d):
        raise _lib.LgcnError("score_topk: the rater was built for another table shape")
    batch_users = min(batch_users, rater.max_users)
    ids = torch.empty((nu, k), dtype=torch.int32, device=dev)
    sc = torch.empty((nu, k), dtype=torch.float32, device=dev)
    fail = torch.empty(nu, dtype=torch.int32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if PROFILE is not None else None
    if ev:
        ev[0].record()
    if rater.prepared_for != (F_item.data_ptr(), F_item._version):
        rater.prepare(F_item)
    if ev:
        ev[1].record()
    for b0 in range(0, nu, batch_users):
        b1 = min(nu, b0 + batch_users)
        # mask_rowptr holds absolute offsets into mask_col, so a batch is a view of both arrays
        mr = mask_rowptr[b0:b1 + 1] if mask_rowptr is not None else None
        rater.topk(F_user, F_item, users[b0:b1], mr, mask_col, k, ids[b0:b1], sc[b0:b1], fail[b0:b1])
    if ev:
        ev[2].record()
        PROFILE.append(("score_tc_prepare", ev[0], ev[1]))
        PROFILE.append(("score_tc_filter_refine", ev[1], ev[2]))
    bad = torch.nonzero(fail).flatten()
    STATS["tc_users"] += nu
    if bad.numel() > 0:                                  # not certified: exact path for those users
        STATS["tc_fallback_users"] += int(bad.numel())
        sub_rp = sub_col = None
        if mask_rowptr is not None:
            sub_rp, sub_col = _sub_csr(mask_rowptr, mask_col, bad)
        i2, s2 = score_topk_exact(F_user, F_item, users[bad].contiguous(), sub_rp, sub_col, k)
        ids[bad] = i2
        sc[bad] = s2
    return ids, sc


@on_device
def eval_metrics(topk_ids, targets, sums=None):
    """sums += [#hits, sum 1/log2(rank+2)] (reference ``main.py:430-438``)."""
    nu, k = topk_ids.shape
    if sums is None:
        sums = torch.zeros(2, dtype=torch.float64, device=topk_ids.device)
    COUNTERS["launches"] += 1
    check(_lib.load().lgcn_eval_metrics(ptr(topk_ids, "i32"), ptr(targets, "i64"), nu, k,
                                        ptr(sums, "f64"), stream_ptr(topk_ids.device)))
    return sums


# ------------------------------------------------------------------------------------------
# autograd glue for the drop-in models (main.py keeps its own loss / optimizer there)
# ------------------------------------------------------------------------------------------
class PropagateFunction(torch.autograd.Function):
    """F = mean_k A^k E0 with E0 = cat(tables) (reference ``models/lightgcn.py:37-59``)."""

    @staticmethod
    def forward(ctx, g, n_layers, *tables):
        e0, alt = _as_block(tables)
        ctx.g, ctx.n_layers = g, n_layers
        ctx.sizes = [t.shape[0] for t in tables]
        return propagate(g, e0, n_layers, alt=alt)

    @staticmethod
    def backward(ctx, grad_f):
        if ctx.g.symmetric is False:
            raise _lib.LgcnError(
                "the backward pass reuses the forward CSR (A^T == A); this adjacency is not "
                "symmetric -- the reference's D^-1/2 A D^-1/2 (main.py:304-331) always is")
        acc = propagate_backward(ctx.g, grad_f.contiguous(), ctx.n_layers)
        return (None, None) + tuple(torch.split(acc, ctx.sizes, dim=0))


def _as_block(tables):
    """(block [N,d], alt): a zero-copy view when the tables are consecutive slices of one
    allocation (models/_packing.py arranges that).  LightGCN_Fusion hands (users, PROJECTED items,
    brands): users and brands still sit in the packed block around the raw item-id rows, so the
    block is used as layer 0 with the projected rows as its override ``alt = (rows, first_row)``
    -- no concat (reference ``lightgcn_fusion.py:52`` builds one every forward).  Anything else is
    concatenated like the reference does (``lightgcn.py:40``)."""
    t0 = tables[0]
    d = t0.shape[1]
    ok = all(t.is_contiguous() and t.shape[1] == d and t.dtype == torch.float32 for t in tables)
    n = sum(t.shape[0] for t in tables)

    def same_storage(t):
        return t.untyped_storage().data_ptr() == t0.untyped_storage().data_ptr()

    if ok:
        end = t0.data_ptr()
        packed = True
        for t in tables:
            if t.data_ptr() != end or not same_storage(t):
                packed = False
                break
            end += t.numel() * 4
        if packed:
            return t0.detach().as_strided((n, d), (d, 1)), None
        if len(tables) == 3:
            u, mid, b = tables
            gap = (u.numel() + mid.numel()) * 4
            room = t0.untyped_storage().nbytes() - (t0.data_ptr() - t0.untyped_storage().data_ptr())
            if same_storage(b) and b.data_ptr() == u.data_ptr() + gap and room >= n * d * 4:
                return (t0.detach().as_strided((n, d), (d, 1), t0.storage_offset()),
                        (mid.detach(), u.shape[0]))
    return torch.cat([t.detach() for t in tables], dim=0), None


class FusionProjFunction(torch.autograd.Function):
    """H = leaky_relu([E_id | C] W^T + b) (reference ``models/lightgcn_fusion.py:45-49``)."""

    @staticmethod
    def forward(ctx, e_id, content, W, b):
        e_id_c, W_c, b_c = e_id.detach().contiguous(), W.detach().contiguous(), b.detach().contiguous()
        H = fusion_proj_fwd(e_id_c, content, W_c, b_c)
        ctx.save_for_backward(e_id_c, content, W_c, H)
        return H

    @staticmethod
    def backward(ctx, gH):
        e_id, content, W, H = ctx.saved_tensors
        g_eid, gW, gb = fusion_proj_bwd(e_id, content, W, H, gH.contiguous())
        return g_eid, None, gW, gb
