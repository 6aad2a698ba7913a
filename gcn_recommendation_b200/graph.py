"""Normalised adjacency in CSR form, resident in HBM.

Host-side mirror of the reference's adjacency builder (reference ``main.py:283-336``) and the
adapter that turns the ``torch.sparse_coo`` tensor ``main.py`` passes to ``forward``
(``main.py:334-336,495``) into the CSR the kernels use.

Layout in HBM (one graph): ``rowptr`` int32[N+1], ``col`` int32[nnz] (ascending inside a row),
``val`` fp32[nnz] = fl32(fl32(d_r*m)*d_c), plus the long-row plan (rows longer than
``long_row_threshold`` are cut into ``seg_len`` segments, include/lgcn.h).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib

LONG_ROW_THRESHOLD_SMALL = 64       # L2-resident (latency-bound) graphs, see _plan defaults below
SEG_LEN_SMALL = 128
LONG_ROW_THRESHOLD_LARGE = 512
SEG_LEN_LARGE = 512
SMALL_GRAPH_ROWS = 1_000_000
TINY_GRAPH_ROWS = 8192
# graphs up to this many rows get the small-graph plan (chunk order, long-row counters); whether the
# small-graph kernels actually run is the library's decision (lgcn_spmm_launches), the plan is
# 4 bytes per 4 rows
SMALL_PLAN_MAX_ROWS = 1 << 20
CHUNK_ORDER_WINDOW = 2048      # chunks per sorting window on large graphs (4-8 rows each)


class NormAdjCSR:
    """D^-1/2 A D^-1/2 as CSR on one GPU (rows ``[row_begin, row_begin+n_rows)`` of an
    ``n_cols``-node graph; the whole graph when not sharded)."""

    def __init__(self, rowptr, col, val, n_cols, row_begin=0, long_row_threshold=None,
                 seg_len=None, rowptr_host=None):
        self.rowptr, self.col, self.val = rowptr, col, val
        self.n_rows = int(rowptr.numel() - 1)
        self.n_cols = int(n_cols)
        self.row_begin = int(row_begin)
        self.nnz = int(col.numel())
        self.device = rowptr.device
        self._seg_ws = {}
        # True: A == A^T was verified; None: built by from_interactions (symmetric by
        # construction, reference main.py:304-313); False: backward through this CSR is refused
        self.symmetric = None
        if long_row_threshold is None and os.environ.get("LGCN_LONG_ROW_THRESHOLD"):
            long_row_threshold = int(os.environ["LGCN_LONG_ROW_THRESHOLD"])      # tuning hook
            seg_len = int(os.environ.get("LGCN_SEG_LEN", max(long_row_threshold // 2, 8)))
        if long_row_threshold is None:
            # small graphs are latency bound: a launch lasts as long as its longest sequential row,
            # so rows above 64 entries go to the (balanced) segment workers; 128-entry segments keep
            # the number of partials the hottest row's combine has to chain low.  Gowalla-shape step:
            # 128/64 0.564 ms, 64/64 0.458, 32/64 0.454, 64/128 0.429
            # (profiles/r02_gowalla_threshold_sweep.txt)
            small = self.n_rows < SMALL_GRAPH_ROWS
            # large graphs (ring kernel: gathers stay pipelined across row boundaries, so long
            # sequential rows are cheap): 512/512 measured best at the Amazon shape (5.06 ms vs
            # 5.23 ms for 256/128); it also leaves fewer rows on the not-bit-exact segment path
            long_row_threshold = LONG_ROW_THRESHOLD_SMALL if small else LONG_ROW_THRESHOLD_LARGE
            seg_len = seg_len or (SEG_LEN_SMALL if small else SEG_LEN_LARGE)
            if self.n_rows < TINY_GRAPH_ROWS:
                # a few thousand rows are one partial wave: nothing to balance, and the sequential
                # (bit-exact) path keeps rows of up to 128 entries (round-1 plan 128 / 64)
                long_row_threshold, seg_len = 128, 64
        self._plan_long_rows(long_row_threshold, seg_len or SEG_LEN_SMALL, rowptr_host)

    # ---- kernel layout (once per graph) ------------------------------------------------
    def _plan_long_rows(self, threshold, seg_len, rowptr_host=None):
        """Build what ``lgcn_spmm`` reads (include/lgcn.h): short rows as packed {col,val} pairs
        with a flagged rowptr, long rows (more than ``threshold`` entries) moved to their own
        CSR and cut into ``seg_len`` segments.  Host logic decides, device ops move the data."""
        self.long_row_threshold = int(threshold)
        self.seg_len = int(seg_len)
        self.n_long = 0
        self.n_seg = 0
        self.long_row_ids = self.long_seg_ptr = self.long_rowptr = self.long_colval = None
        dev = self.device
        rp = rowptr_host if rowptr_host is not None else self.rowptr.cpu().numpy()
        deg = np.diff(rp.astype(np.int64))
        long_ids = (np.nonzero(deg > threshold)[0] if threshold > 0 else np.zeros(0, np.int64))
        packed = torch.stack([self.col, self.val.view(torch.int32)], dim=1)     # [nnz, 2] int32
        self.long_done = None
        self._chunk_orders = {}
        if len(long_ids) == 0:
            self.long_row_threshold = 0 if threshold <= 0 else self.long_row_threshold
            self.rowptr_flagged = self.rowptr
            self.colval = packed.contiguous()
            self._plan_small_graph(deg)
            return
        is_long = np.zeros(self.n_rows, bool)
        is_long[long_ids] = True
        short_deg = np.where(is_long, 0, deg)
        rp_short = np.zeros(self.n_rows + 1, np.int64)
        np.cumsum(short_deg, out=rp_short[1:])
        flagged = rp_short.astype(np.uint32)
        flagged[:-1] |= (is_long.astype(np.uint32) << np.uint32(31))
        long_deg = deg[long_ids]
        long_rp = np.zeros(len(long_ids) + 1, np.int32)
        np.cumsum(long_deg, out=long_rp[1:])
        nseg = (long_deg + seg_len - 1) // seg_len
        seg_ptr = np.zeros(len(long_ids) + 1, np.int32)
        np.cumsum(nseg, out=seg_ptr[1:])
        # entry mask: which entries belong to long rows (device, from the row flags)
        long_flag = torch.from_numpy(is_long).to(dev)
        entry_long = torch.repeat_interleave(long_flag, torch.from_numpy(deg).to(dev))
        self.colval = packed[~entry_long].contiguous()
        self.long_colval = packed[entry_long].contiguous()
        self.rowptr_flagged = torch.from_numpy(flagged.view(np.int32)).to(dev)
        self.n_long = int(len(long_ids))
        self.n_seg = int(seg_ptr[-1])
        self.long_row_ids = torch.from_numpy(long_ids.astype(np.int32)).to(dev)
        self.long_rowptr = torch.from_numpy(long_rp).to(dev)
        self.long_seg_ptr = torch.from_numpy(seg_ptr).to(dev)
        self._plan_small_graph(short_deg)

    def _plan_small_graph(self, short_deg=None):
        """Long-row counters of the small-graph (L2-resident) kernels, include/lgcn.h ``long_done``:
        the worker delivering a long row's last segment combines it, no combine launch is needed.
        Results do not depend on it (tests run with and without: ``LGCN_NO_SMALL_PLAN=1``)."""
        self._chunk_orders = {}
        plan = not os.environ.get("LGCN_NO_SMALL_PLAN")            # decided when the graph is built
        self._order_enabled = plan and not os.environ.get("LGCN_NO_CHUNK_ORDER")
        if self.n_rows > SMALL_PLAN_MAX_ROWS or not plan:
            return
        if self.n_long > 0:
            self.long_done = torch.zeros(self.n_long, dtype=torch.int32, device=self.device)

    def chunk_order_for(self, rows_per_chunk, windowed):
        """``lgcn_spmm_args.chunk_order`` for chunks of ``rows_per_chunk`` rows (include/lgcn.h):
        chunks of similar entry count next to each other, so that the workers sharing a warp (tables
        narrower than 128 floats) walk equally long streams.  Small graphs: one global descending
        sort (longest work first).  Large graphs (``windowed``): descending inside windows of
        CHUNK_ORDER_WINDOW chunks, so that entries, outputs and epilogue operands still stream
        through HBM window by window.  Pure device integer work, once per (graph, chunk size)."""
        if not getattr(self, "_order_enabled", False):
            return None
        key = (int(rows_per_chunk), bool(windowed))
        orders = self.__dict__.setdefault("_chunk_orders", {})
        if key not in orders:
            R = key[0]
            rp = self.rowptr_flagged.to(torch.int64) & 0x7fffffff       # short-row prefix sums
            n_chunks = (self.n_rows + R - 1) // R
            deg = torch.zeros(n_chunks * R, dtype=torch.int64, device=self.device)
            deg[:self.n_rows] = rp[1:] - rp[:-1]
            lens = deg.view(n_chunks, R).sum(1)
            if not windowed:
                order = torch.sort(lens, descending=True, stable=True).indices
            else:
                W = int(os.environ.get("LGCN_CHUNK_ORDER_WINDOW", CHUNK_ORDER_WINDOW))
                n_win = (n_chunks + W - 1) // W
                padded = torch.full((n_win * W,), -1, dtype=torch.int64, device=self.device)
                padded[:n_chunks] = lens
                idx = torch.sort(padded.view(n_win, W), dim=1, descending=True, stable=True).indices
                idx = (idx + torch.arange(n_win, device=self.device)[:, None] * W).reshape(-1)
                order = idx[idx < n_chunks]
            orders[key] = order.to(torch.int32).contiguous()
        return orders[key]

    # ---- L2 residency classes of the gathered columns (include/lgcn.h, LGCN_COL_*) ----------
    def mark_hot_columns(self, n_hot, col_degree=None):
        """Classify the columns of the short-row entries for the kernels' per-gather L2 policy:
        the ``n_hot`` highest-degree nodes become HOT (bit 31 of ``colval[:,0]``: gathered rows
        kept with evict_last), nodes of degree 1 -- gathered once per launch -- ONCE (bit 30:
        evict_first).  The indices themselves (low 30 bits) are untouched; ``n_hot == 0`` clears
        the classes.  Pure device integer work on the plan, done once per (graph, table width)."""
        n_hot = max(0, min(int(n_hot), self.n_cols))
        if getattr(self, "n_hot", 0) == n_hot and (n_hot == 0 or col_degree is None):
            return
        if self.n_cols > (1 << 30):
            raise ValueError("column classes need n_cols <= 2^30")
        if col_degree is None:
            col_degree = getattr(self, "col_degree", None)
        if col_degree is None:
            if self.n_rows != self.n_cols:
                raise ValueError("a row shard needs the column degrees of the whole graph")
            col_degree = (self.rowptr[1:] - self.rowptr[:-1])        # symmetric: column = row degree
        self.col_degree = col_degree
        cls_bits = torch.zeros(self.n_cols, dtype=torch.int32, device=self.device)
        if n_hot > 0:
            cls_bits[col_degree == 1] = 0x40000000
            hot = torch.topk(col_degree, n_hot, sorted=False).indices
            hot = hot[col_degree[hot] > 1]
            cls_bits[hot] = -0x80000000
        cols = self.colval[:, 0] & 0x3fffffff
        self.colval[:, 0] = cols | cls_bits[cols.long()]
        self.n_hot = n_hot

    def seg_ws(self, d):
        if self.n_seg == 0:
            return None
        ws = self._seg_ws.get(d)
        if ws is None:
            ws = torch.empty((self.n_seg, d), dtype=torch.float32, device=self.device)
            self._seg_ws[d] = ws
        return ws

    def degrees_host(self):
        return np.diff(self.rowptr.cpu().numpy().astype(np.int64))

    # ---- builders ------------------------------------------------------------------------
    @classmethod
    def from_interactions(cls, train_user, train_item, num_users, num_items, num_brands, device,
                          item_brand=None, **kw):
        """Build from the training interactions exactly as reference ``main.py:283-336`` does:
        symmetric COO of ones over users|items|brands (``:304-313``), duplicates summed by the
        CSR conversion, ``d = np.power(rowsum, -0.5)`` in fp32 with inf -> 0 (``:326-329``) --
        the SAME numpy call on the host so the weights are bit-equal -- and the product
        ``fl32(fl32(d_r*m)*d_c)`` (``:330-331``) formed on the GPU by ``lgcn_edge_weights``."""
        U, I, B = int(num_users), int(num_items), int(num_brands)
        N = U + I + B
        if isinstance(train_user, torch.Tensor) and train_user.is_cuda:
            return cls._from_interactions_device(train_user, train_item, U, I, B, item_brand, **kw)
        u = np.ascontiguousarray(train_user, dtype=np.int64)
        it = np.ascontiguousarray(train_item, dtype=np.int64) + U
        if len(u) and (u.min() < 0 or u.max() >= U or it.min() < U or it.max() >= U + I):
            raise ValueError("interaction index out of range")
        rows, cols = [u, it], [it, u]
        if item_brand is not None:
            ib_i = np.ascontiguousarray(item_brand[0], dtype=np.int64) + U
            ib_b = np.ascontiguousarray(item_brand[1], dtype=np.int64) + U + I
            rows += [ib_i, ib_b]
            cols += [ib_b, ib_i]
        key = np.concatenate(rows) * N + np.concatenate(cols)
        del rows, cols
        key.sort()
        if len(key):
            first = np.ones(len(key), dtype=bool)
            first[1:] = key[1:] != key[:-1]
            ukey = key[first]
            pos = np.flatnonzero(first)
            mult = np.diff(np.append(pos, len(key))).astype(np.float32)
        else:
            ukey, mult = key, np.zeros(0, np.float32)
        del key
        r = ukey // N
        c = (ukey - r * N).astype(np.int32)
        counts = np.bincount(r, minlength=N)
        rowptr = np.zeros(N + 1, np.int64)
        np.cumsum(counts, out=rowptr[1:])
        if rowptr[-1] >= 2 ** 31:
            raise ValueError("nnz does not fit int32")
        unit = bool(len(mult) == 0 or mult.max() == 1.0)
        deg = (counts if unit else np.bincount(r, weights=mult, minlength=N)).astype(np.float32)
        with np.errstate(divide="ignore"):
            dinv = np.power(deg, np.float32(-0.5)).astype(np.float32)   # main.py:328
        dinv[np.isinf(dinv)] = 0.0                                       # main.py:329
        rp32 = rowptr.astype(np.int32)
        t_rowptr = torch.from_numpy(rp32).to(device)
        t_col = torch.from_numpy(c).to(device)
        t_dinv = torch.from_numpy(dinv).to(device)
        t_mult = None if unit else torch.from_numpy(mult).to(device)
        t_val = torch.empty(len(c), dtype=torch.float32, device=device)
        lib = _lib.load()
        _lib.check(lib.lgcn_edge_weights(_lib.ptr(t_rowptr, "i32"), _lib.ptr(t_col, "i32"),
                                         _lib.ptr(t_dinv), _lib.ptr(t_mult, allow_none=True),
                                         _lib.ptr(t_val), N, _lib.stream_ptr(device)))
        g = cls(t_rowptr, t_col, t_val, N, rowptr_host=rp32, **kw)
        g.dinv = t_dinv
        return g

    @classmethod
    def _from_interactions_device(cls, train_user, train_item, U, I, B, item_brand=None, **kw):
        """Same construction with the edge list already in HBM: the (row, col) sort / dedupe /
        degree count run as torch device ops (integer work, bit-exact); ``d = np.power(deg,
        -0.5)`` is still the host numpy call of reference ``main.py:328`` (O(N))."""
        dev = train_user.device
        N = U + I + B
        u = train_user.to(torch.int64)
        it = train_item.to(torch.int64) + U
        parts = [u * N + it, it * N + u]
        if item_brand is not None:
            ib_i = torch.as_tensor(item_brand[0], device=dev).to(torch.int64) + U
            ib_b = torch.as_tensor(item_brand[1], device=dev).to(torch.int64) + U + I
            parts += [ib_i * N + ib_b, ib_b * N + ib_i]
        key = torch.sort(torch.cat(parts)).values
        del parts
        ukey, mult = torch.unique_consecutive(key, return_counts=True)
        del key
        r = torch.div(ukey, N, rounding_mode="floor")
        c = (ukey - r * N).to(torch.int32)
        counts = torch.bincount(r, minlength=N)
        rowptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=rowptr[1:])
        if int(rowptr[-1].item()) >= 2 ** 31:
            raise ValueError("nnz does not fit int32")
        unit = bool(mult.numel() == 0 or int(mult.max().item()) == 1)
        deg = counts if unit else torch.bincount(r, weights=mult.to(torch.float64), minlength=N)
        del r, ukey
        deg_h = deg.cpu().numpy().astype(np.float32)
        with np.errstate(divide="ignore"):
            dinv = np.power(deg_h, np.float32(-0.5)).astype(np.float32)   # main.py:328
        dinv[np.isinf(dinv)] = 0.0                                         # main.py:329
        t_rowptr = rowptr.to(torch.int32)
        t_dinv = torch.from_numpy(dinv).to(dev)
        t_mult = None if unit else mult.to(torch.float32)
        t_val = torch.empty(c.numel(), dtype=torch.float32, device=dev)
        lib = _lib.load()
        _lib.check(lib.lgcn_edge_weights(_lib.ptr(t_rowptr, "i32"), _lib.ptr(c, "i32"),
                                         _lib.ptr(t_dinv), _lib.ptr(t_mult, allow_none=True),
                                         _lib.ptr(t_val), N, _lib.stream_ptr(dev)))
        g = cls(t_rowptr, c, t_val, N, **kw)
        g.dinv = t_dinv
        return g

    @classmethod
    def from_sparse_coo(cls, adj_mat, **kw):
        """Adapter for the uncoalesced fp32 ``torch.sparse_coo`` tensor of reference
        ``main.py:334-336``.  It is row-major sorted with unique entries (it came out of a scipy
        CSR); that is verified on the device, and anything else is coalesced (sorted, duplicates
        summed -- what ``torch.sparse.mm`` itself would do) first."""
        if adj_mat.layout != torch.sparse_coo:
            raise TypeError("adj_mat must be a torch.sparse_coo tensor")
        if not adj_mat.is_cuda:
            raise _lib.LgcnError("adj_mat must live on a CUDA device (no CPU fallback)")
        N = adj_mat.shape[0]
        lib = _lib.load()
        for attempt in range(2):
            idx = adj_mat._indices()
            vals = adj_mat._values()
            if vals.dtype != torch.float32:
                raise TypeError("adjacency values must be float32")
            row = idx[0].contiguous()
            colin = idx[1].contiguous()
            nnz = int(vals.numel())
            rowptr = torch.empty(N + 1, dtype=torch.int32, device=adj_mat.device)
            col = torch.empty(nnz, dtype=torch.int32, device=adj_mat.device)
            status = torch.zeros(1, dtype=torch.int32, device=adj_mat.device)
            with torch.cuda.device(adj_mat.device):
                _lib.check(lib.lgcn_csr_from_sorted_coo(_lib.ptr(row, "i64"), _lib.ptr(colin, "i64"), nnz,
                                                        N, int(adj_mat.shape[1]), _lib.ptr(rowptr, "i32"),
                                                        _lib.ptr(col, "i32"), _lib.ptr(status, "i32"),
                                                        _lib.stream_ptr(adj_mat.device)))
            if int(status.item()) == 0:
                g = cls(rowptr, col, vals.contiguous().clone(), adj_mat.shape[1], **kw)
                g.symmetric = _is_symmetric(row, colin, g.val, N, adj_mat.shape[1])
                return g
            if attempt == 0:
                adj_mat = adj_mat.coalesce()
        raise _lib.LgcnError("adjacency indices are out of range")

    def row_shard(self, row_begin, row_end):
        """CSR of rows [row_begin, row_end) (columns stay global) for row-sharded propagation."""
        rp = self.rowptr[row_begin:row_end + 1]
        e0, e1 = int(rp[0].item()), int(rp[-1].item())
        g = NormAdjCSR((rp - e0).contiguous(), self.col[e0:e1].contiguous(),
                       self.val[e0:e1].contiguous(), self.n_cols, row_begin=row_begin,
                       long_row_threshold=self.long_row_threshold or 0, seg_len=self.seg_len)
        if self.n_rows == self.n_cols:
            g.col_degree = self.rowptr[1:] - self.rowptr[:-1]
        return g


def _is_symmetric(row, col, val, n_rows, n_cols):
    """A == A^T for a strictly (row, col)-sorted COO: the transposed keys, sorted, must be the same
    key list with the same values.  Device integer work, once per adjacency (the backward pass
    reuses the forward CSR, which is only valid for a symmetric matrix)."""
    if n_rows != n_cols:
        return False
    key_t, order = torch.sort(col * n_rows + row)
    if not torch.equal(key_t, row * n_rows + col):
        return False
    # D^-1/2 A D^-1/2 rounds (d_r*m)*d_c and (d_c*m)*d_r separately: up to 1 ulp apart
    return bool(torch.allclose(val[order], val, rtol=1e-6, atol=0.0))


_COO_CACHE = {}


def csr_for(adj_mat):
    """CSR for the ``adj_mat`` object ``main.py`` passes on every call (``main.py:495,413``),
    converted once and cached on the identity of its index / value storage."""
    key = (adj_mat._indices().data_ptr(), adj_mat._values().data_ptr(), int(adj_mat._nnz()),
           tuple(adj_mat.shape), str(adj_mat.device))
    g = _COO_CACHE.get(key)
    if g is None:
        if len(_COO_CACHE) > 8:
            _COO_CACHE.clear()
        g = NormAdjCSR.from_sparse_coo(adj_mat)
        g._keepalive = adj_mat          # the key is only valid while the tensor lives
        _COO_CACHE[key] = g
    return g
