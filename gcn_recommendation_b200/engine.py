"""Native LightGCN trainer / evaluator: the throughput vehicle (mode 2).

One training step is what reference ``main.py:488-531`` does per batch -- full K-layer
propagation, six row gathers, ``bpr_loss_reg``, backward, dense Adam -- restated as

  forward   K launches of the CSR SpMM, layer mean fused into the last        (a2)
  bpr       one fused gather/dot/log-sigmoid/L2/scatter-add launch            (a4)
  backward  K-1 Horner hops  acc <- g' + A acc  with the same SpMM kernel     (a3)
  adam      fused into the epilogue of the K-th hop (dense grad never stored) (a5)

on buffers that are allocated once, so a step is CUDA-graph capturable and nothing syncs
with the host unless the caller reads the loss.  Evaluation is reference ``main.py:404-439``:
one propagation, then fused score + mask + top-k and the hit / NDCG sums on the device.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from . import ops
from ._lib import LgcnError

_CHECK_ALWAYS = os.environ.get("LGCN_CHECK_INDICES", "0") == "1"


def build_mask_csr(eval_users, train_user, train_item, num_users, device=None):
    """Per-evaluated-user ascending train-item lists as CSR over the position in ``eval_users``
    (what ``train_df.groupby('user_idx')['item_idx'].apply(list)`` feeds the mask loop of
    reference ``main.py:407,422-424``).  Returns int64 rowptr / int32 col tensors.  Host numpy
    arrays are processed on the host; CUDA tensors stay on the device (sort + searchsorted)."""
    if isinstance(train_user, torch.Tensor) and train_user.is_cuda:
        dev = train_user.device
        tu, ti = train_user.to(torch.int64), train_item.to(torch.int64)
        n_items = int(ti.max().item()) + 1 if ti.numel() else 1
        key = torch.sort(tu * n_items + ti).values
        start = torch.searchsorted(key, torch.arange(num_users + 1, device=dev, dtype=torch.int64) * n_items)
        eu = torch.as_tensor(eval_users, device=dev).to(torch.int64)
        rp, cc = ops._sub_csr(start, (key % n_items).to(torch.int32), eu)
        return rp, cc
    tu = np.ascontiguousarray(train_user, np.int64)
    ti = np.ascontiguousarray(train_item, np.int64)
    order = np.lexsort((ti, tu))
    tu, ti = tu[order], ti[order]
    start = np.searchsorted(tu, np.arange(num_users + 1, dtype=np.int64))
    eu = np.ascontiguousarray(eval_users, np.int64)
    cnt = start[eu + 1] - start[eu]
    rowptr = np.zeros(len(eu) + 1, np.int64)
    np.cumsum(cnt, out=rowptr[1:])
    src = np.repeat(start[eu] - rowptr[:-1], cnt) + np.arange(rowptr[-1], dtype=np.int64)
    col = ti[src].astype(np.int32)
    rp, cc = torch.from_numpy(rowptr), torch.from_numpy(col)
    if device is not None:
        rp, cc = rp.to(device), cc.to(device)
    return rp, cc


def mask_csr_from_graph(g, eval_users, num_users):
    """The same mask lists read off the graph itself: rows [0, U) of the adjacency ARE the users'
    train items (columns U + item, ascending), so validation-time masks (mask = training
    interactions, reference ``main.py:407,545``) need no second sort.  Device ops only."""
    eu = torch.as_tensor(eval_users, device=g.device).to(torch.int64)
    return ops._sub_csr(g.rowptr.to(torch.int64), g.col - int(num_users), eu)


def xavier_uniform_table(rows_list, d, device, generator=None):
    """users | items | brands block, each part ``nn.init.xavier_uniform_`` like reference
    ``models/lightgcn.py:27-31`` (bound = sqrt(6/(rows+d)))."""
    n = sum(rows_list)
    block = torch.empty((n, d), dtype=torch.float32, device=device)
    r = 0
    for rows in rows_list:
        bound = math.sqrt(6.0 / (rows + d)) if rows > 0 else 0.0
        block[r:r + rows].uniform_(-bound, bound, generator=generator)
        r += rows
    return block


class LightGCNEngine:
    """Fused LightGCN step + full-rank evaluation on one GPU.

    ``table``: the [N,d] users|items|brands parameter block (shared with the drop-in module
    when created through ``LightGCN.engine``).  ``fusion``: optional dict(content=[I,c],
    weight=[d,d+c], bias=[d]) enabling the ``LightGCN_Fusion`` item block (reference
    ``models/lightgcn_fusion.py:45-49``); ``table`` then holds the raw id embeddings.
    """

    def __init__(self, g, num_users, num_items, num_brands, n_layers, table, lr=1e-3,
                 weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, fusion=None, batch_size=2048,
                 item_to_brand=None, brand_loss_weight=0.0):
        self.g = g
        self.U, self.I, self.B = int(num_users), int(num_items), int(num_brands)
        self.N = self.U + self.I + self.B
        self.K = int(n_layers)
        if tuple(table.shape[:1]) != (self.N,) or not table.is_contiguous():
            raise LgcnError("table must be a contiguous [num_users+num_items+num_brands, d] block")
        if g.n_rows != self.N or g.n_cols != self.N:
            raise LgcnError("graph size does not match the table")
        self.P = table
        self.d = int(table.shape[1])
        self.dev = table.device
        self.lr, self.lam, self.betas, self.eps = float(lr), float(weight_decay), betas, float(eps)
        new = lambda: torch.empty((self.N, self.d), dtype=torch.float32, device=self.dev)  # noqa: E731
        self.m = torch.zeros_like(self.P)
        self.v = torch.zeros_like(self.P)
        self.F = new()
        # E_1..E_{K-1}, reused as the Horner ping-pong of the K-1 ADD hops (never F: a later
        # rate_topk(propagate=False) reads F)
        self.work = [new() for _ in range(max(self.K - 1, 2 if self.K > 1 else 0))]
        self.G1 = torch.zeros_like(self.P)          # g/(K+1): addend of every Horner hop
        self.G2 = torch.zeros_like(self.P)          # regulariser grad (+ g/(K+1) w/o fusion)
        # 1 = row received a gradient this step (G1/G2 are zero elsewhere): lets the backward
        # hops skip the gathers / addend reads of all-zero rows
        self.rowflag = torch.zeros(self.N + 32, dtype=torch.uint8, device=self.dev)   # padded
        # rows of the first hop's output that can be non-zero (written by the kernel): the second
        # hop gathers only those, and the first hop does not write the others at all
        self.rowflag2 = torch.zeros(self.N + 32, dtype=torch.uint8, device=self.dev)
        self.zero_row = torch.zeros(256, dtype=torch.float32, device=self.dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.adam_scalars = torch.zeros(2, dtype=torch.float32, device=self.dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self.idx_status = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.bs = int(batch_size)
        # optional brand / author BPR term (reference main.py:382-391): item -> brand lookup + weight
        self.brand_w = float(brand_loss_weight)
        self.item_brand = None
        if self.brand_w != 0.0:
            if item_to_brand is None or self.B <= 0:
                raise LgcnError("brand_loss_weight needs item_to_brand and num_brands > 0")
            self.item_brand = torch.as_tensor(item_to_brand, device=self.dev).to(torch.int64).contiguous()
            if self.item_brand.numel() != self.I:
                raise LgcnError("item_to_brand must have one entry per item")
            if self.item_brand.numel() and (int(self.item_brand.min()) < 0 or int(self.item_brand.max()) >= self.B):
                raise IndexError("item_to_brand index out of range")
            self.loss_brand = torch.zeros(1, dtype=torch.float32, device=self.dev)
        # Sparse first-hop output + flagged second hop: pays when the tables stream from HBM
        # (Amazon shape: 4.6 + 5.2 -> 1.8 + 3.5 ms); on L2-resident graphs the extra flag lookups
        # only lengthen a latency-bound kernel (Gowalla shape: 0.66 -> 0.78 ms per step).
        # LGCN_SPARSE_HOPS=0/1 forces it (A/B measurements, tests).
        env = os.environ.get("LGCN_SPARSE_HOPS")
        self.sparse_hops = (self.N * self.d * 4 > ops.L2_STREAM_BYTES) if env is None else env != "0"
        self._alloc_batch(self.bs)
        self.fusion = None
        if fusion is not None:
            self._init_fusion(fusion)
        self._graph = None
        self._tc = None                             # ops.TcRater of the evaluation sweep
        self._idx_checked_eval = None
        self.launches_per_step = self._count_launches()

    # ---- setup ---------------------------------------------------------------------------
    def _alloc_batch(self, bs):
        self.bs = bs
        self._checked_bs = -1
        self.b_users = torch.zeros(bs, dtype=torch.int64, device=self.dev)
        self.b_pos = torch.zeros(bs, dtype=torch.int64, device=self.dev)
        self.b_neg = torch.zeros(bs, dtype=torch.int64, device=self.dev)
        self.sample_ws = torch.empty(2 * bs, dtype=torch.float32, device=self.dev)
        self.b_bpos = torch.zeros(bs, dtype=torch.int64, device=self.dev)     # brands of pos / neg items
        self.b_bneg = torch.zeros(bs, dtype=torch.int64, device=self.dev)
        self._graph = None

    def _init_fusion(self, f):
        C, W, b = f["content"], f["weight"], f["bias"]
        if C.shape[0] != self.I or W.shape != (self.d, self.d + C.shape[1]) or b.shape != (self.d,):
            raise LgcnError("fusion: shape mismatch")
        z = torch.zeros_like
        self.fusion = dict(C=C.contiguous(), W=W, b=b, mW=z(W), vW=z(W), mb=z(b), vb=z(b),
                           gW=z(W), gb=z(b),
                           g_eid=torch.empty((self.I, self.d), dtype=torch.float32, device=self.dev))
        # projected item rows H (layer 0 = the parameter table with its item rows overridden by H:
        # lgcn_spmm's x_alt, so the concatenated table of reference lightgcn_fusion.py:52 is never
        # built) and dL/dH, written by the ADAM hop for the rows it skips
        self.H = torch.empty((self.I, self.d), dtype=torch.float32, device=self.dev)
        self.gH = torch.empty((self.I, self.d), dtype=torch.float32, device=self.dev)

    def _count_launches(self):
        """Kernels of THIS library launched per training step (for bench.py's gpu_launches)."""
        per_spmm = ops.spmm_launches(self.g, self.d)
        n = 2 * self.K * per_spmm + 2 + 1 + 1          # spmm fwd+bwd, bpr(+reduce), tick, zero
        if self.brand_w != 0.0:
            n += 2 + 1                                 # brand term (+reduce), its zero_rows
        if self.fusion is not None:
            n += 1 + 2 + 3                             # proj fwd, proj bwd (2), Adam of item ids / W / b
        return n

    # ---- pieces ----------------------------------------------------------------------------
    def layer0(self):
        """Input of the propagation as (table, alt): the parameter table itself, and with the
        fusion block ``alt = (H, U)`` -- its item rows are the projected rows H (reference
        ``models/lightgcn_fusion.py:45-52``), read in place of the raw id rows by the kernels."""
        if self.fusion is None:
            return self.P, None
        U, I = self.U, self.I
        f = self.fusion
        ops.fusion_proj_fwd(self.P[U:U + I], f["C"], f["W"], f["b"], out=self.H)
        return self.P, (self.H, U)

    def _fusion_backward(self, gH):
        """Gradients of the fused item block from dL/dH (``gH`` [I, d]): fills ``g_eid`` (id
        embeddings), ``gW`` and ``gb`` (autograd of reference ``models/lightgcn_fusion.py:45-49``)."""
        U, I = self.U, self.I
        f = self.fusion
        f["gW"].zero_()
        f["gb"].zero_()
        ops.fusion_proj_bwd(self.P[U:U + I], f["C"], f["W"], self.H, gH,
                            g_eid=f["g_eid"], gW=f["gW"], gb=f["gb"])

    def propagate(self):
        """F = mean_k A^k E0 (reference ``models/lightgcn.py:44-54``); returns the [N,d] table."""
        if self._tc is not None:
            self._tc.prepared_for = None            # the bf16 item tiles belong to the old table
        x, alt = self.layer0()
        return ops.propagate(self.g, x, self.K, out=self.F, work=self.work, alt=alt)

    def forward(self):
        """The reference's 5-tuple (``models/lightgcn.py:81``) as views, no autograd."""
        F = self.propagate()
        U, I = self.U, self.I
        return F[:U], F[U:U + I], F[U + I:], self.P[:U], self.P[U:U + I]

    def _bpr_term(self, F, pos, neg, item_offset, lam, scale, gp_includes_gf, loss_out):
        """One BPR term over (b_users, pos, neg) rows at ``item_offset``: loss into ``loss_out``,
        ``scale * dL/dF`` scattered into G1 and the regulariser gradient (+ the same dL/dF when
        ``gp_includes_gf``) into G2, touched rows flagged."""
        ops.bpr_fused(F, self.P, self.b_users, pos, neg, item_offset, lam, grad_scale=scale,
                      gF=self.G1, gP=self.G2, gp_includes_gf=gp_includes_gf, sample_ws=self.sample_ws,
                      loss_out=loss_out, rowflag=self.rowflag)

    def _bpr(self, F, gp_includes_gf):
        """Loss + scatter of dL/dF (scaled by 1/(K+1)) into G1 and of the regulariser gradient
        into G2 for the staged batch (reference ``main.py:496-497,515-525``); with
        ``brand_loss_weight`` also the brand term of ``main.py:382-391,401`` (same kernel over the
        brand rows of the batch's items, no regulariser, weighted)."""
        scale = 1.0 / (self.K + 1)
        self._bpr_term(F, self.b_pos, self.b_neg, self.U, self.lam, scale, gp_includes_gf, self.loss)
        if self.brand_w != 0.0:
            torch.index_select(self.item_brand, 0, self.b_pos, out=self.b_bpos)
            torch.index_select(self.item_brand, 0, self.b_neg, out=self.b_bneg)
            self._bpr_term(F, self.b_bpos, self.b_bneg, self.U + self.I, 0.0, scale * self.brand_w,
                           gp_includes_gf, self.loss_brand)
            self.loss.add_(self.loss_brand, alpha=self.brand_w)

    def _step_body(self):
        g, K, U, I = self.g, self.K, self.U, self.I
        u, p, n = self.b_users, self.b_pos, self.b_neg
        F = self.propagate()
        nofus = self.fusion is None
        self._bpr(F, nofus)
        ops.adam_tick(self.step_dev, self.adam_scalars, self.lr, self.betas)
        acc = self.G1
        rf, zr = self.rowflag, self.zero_row
        rf2 = self.rowflag2 if (self.sparse_hops and K - 1 >= 2) else None
        for k in range(K - 1):
            # hop 0 gathers g' itself (<= 3*batch non-zero rows): flagged gathers, and only the
            # rows next to the batch's nodes are written (flags in rf2); hop 1 gathers under rf2
            # (still ~80 % zero rows at the Amazon shape).  The addend of every hop is g' too: its
            # all-zero rows are not read (1 ms per hop at the Amazon shape)
            acc = ops.spmm(g, acc, out=self.work[k % 2], addend=self.G1,
                           x_rowflag=rf if k == 0 else (rf2 if k == 1 else None), addend_rowflag=rf,
                           zero_row=zr, y_rowflag=rf2 if k == 0 else None)
        if nofus:
            ops.spmm_adam(g, acc, self.P, self.m, self.v, self.adam_scalars, addend=self.G2,
                          betas=self.betas, eps=self.eps, addend_rowflag=rf, zero_row=zr)
        else:
            # the K-th hop updates the user / brand rows (grad = g' + A acc + reg) and hands
            # dL/dH = g' + A acc of the item rows to the projection's backward; the item-id rows,
            # W and b then take their own (small) Adam launches
            f = self.fusion
            ops.spmm_adam(g, acc, self.P, self.m, self.v, self.adam_scalars, addend=self.G1,
                          addend2=self.G2, betas=self.betas, eps=self.eps, addend_rowflag=rf,
                          zero_row=zr, skip=(self.gH, U))
            self._fusion_backward(self.gH)
            kw = dict(betas=self.betas, eps=self.eps)
            sc = self.adam_scalars
            ops.adam(self.P[U:U + I], f["g_eid"], self.m[U:U + I], self.v[U:U + I], sc,
                     g1=self.G2[U:U + I], **kw)
            ops.adam(f["W"], f["gW"], f["mW"], f["vW"], sc, **kw)
            ops.adam(f["b"], f["gb"], f["mb"], f["vb"], sc, **kw)
        ops.zero_rows(self.G1, self.G2, u, p, n, U, rowflag=self.rowflag)
        if self.brand_w != 0.0:
            ops.zero_rows(self.G1, self.G2, u, self.b_bpos, self.b_bneg, U + I, rowflag=self.rowflag)

    # ---- public --------------------------------------------------------------------------
    def capture(self):
        """Capture one training step (static batch buffers) in a CUDA graph."""
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        saved = self._snapshot()
        with torch.cuda.stream(s):
            self._step_body()                       # warm-up outside capture
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self._restore(saved)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._step_body()
        self._restore(saved)                        # capture does not execute, but be explicit
        self._graph = graph
        return graph

    def _snapshot(self):
        st = [self.P.clone(), self.m.clone(), self.v.clone(), self.step_dev.clone()]
        if self.fusion is not None:
            f = self.fusion
            st += [f[k].clone() for k in ("W", "b", "mW", "vW", "mb", "vb")]
        return st

    def _restore(self, st):
        self.P.copy_(st[0]); self.m.copy_(st[1]); self.v.copy_(st[2]); self.step_dev.copy_(st[3])
        if self.fusion is not None:
            f = self.fusion
            for k, t in zip(("W", "b", "mW", "vW", "mb", "vb"), st[4:]):
                f[k].copy_(t)

    def _check_batch(self, users, pos, neg):
        """Range check of a batch (the reference's gathers raise IndexError, main.py:496-497).
        Costs a device->host read, so it runs for the first batch of every new batch size and on
        every batch with LGCN_CHECK_INDICES=1."""
        self.idx_status.zero_()
        ops.check_indices(self.idx_status, (users, 0, self.U), (pos, 0, self.I), (neg, 0, self.I))
        if int(self.idx_status.item()) != 0:
            raise IndexError(f"{int(self.idx_status.item())} batch indices are out of range "
                             f"(users < {self.U}, items < {self.I})")

    def bpr_step(self, users, pos, neg, use_graph=True):
        """One fused training step on a (users, pos, neg) int64 batch (host or device tensors).
        Returns the device loss tensor [1] (read it with ``.item()`` only when needed; reference
        ``main.py:528`` syncs every step)."""
        bs = users.numel()
        if not (pos.numel() == bs and neg.numel() == bs):
            raise LgcnError("users, pos and neg must have the same length")
        if bs > self.bs or (bs < self.bs and self._graph is None) or bs == 0:
            self._alloc_batch(bs)
        if bs < self.bs:
            # a shorter (tail) batch while a step is captured for the full size: run it eagerly on
            # views of the staging buffers, the captured graph stays valid
            return self._tail_step(users, pos, neg, bs)
        self.b_users.copy_(users, non_blocking=True)
        self.b_pos.copy_(pos, non_blocking=True)
        self.b_neg.copy_(neg, non_blocking=True)
        if self._checked_bs != bs or _CHECK_ALWAYS:
            self._check_batch(self.b_users, self.b_pos, self.b_neg)
            self._checked_bs = bs
        if use_graph:
            if self._graph is None:
                self.capture()
            self._graph.replay()
        else:
            self._step_body()
        return self.loss

    _BATCH_BUFFERS = ("b_users", "b_pos", "b_neg", "b_bpos", "b_bneg")

    def _tail_step(self, users, pos, neg, bs):
        full = {k: getattr(self, k) for k in self._BATCH_BUFFERS + ("sample_ws",) + self._extra_batch_buffers()}
        try:
            for k, t in full.items():
                per = t.numel() // self.bs
                setattr(self, k, t[:bs * per])
            self.b_users.copy_(users, non_blocking=True)
            self.b_pos.copy_(pos, non_blocking=True)
            self.b_neg.copy_(neg, non_blocking=True)
            self._check_batch(self.b_users, self.b_pos, self.b_neg)
            self._step_body()
        finally:
            for k, t in full.items():
                setattr(self, k, t)
        return self.loss

    def _extra_batch_buffers(self):
        return ()

    def train_steps(self, n_steps, seed=42, use_graph=False):
        """``n_steps`` training steps on batches drawn by the device-side sampler (replaces the
        reference's DataLoader, ``main.py:462-464,488``): nothing crosses the PCIe bus.  Returns
        the device loss of the last step.

        An epoch visits every UNIQUE (user, item) training pair once (``rowptr[U]`` entries of the
        CSR); the reference's DataLoader visits ``len(train_df)`` rows, i.e. a repeated interaction
        as many times as it occurs in train.parquet (``main.py:349-363``).  The two agree on
        deduplicated data, which is what the reference's prepare_data scripts write."""
        if not hasattr(self, "sampler_state"):
            self.sampler_state = torch.zeros(2, dtype=torch.int64, device=self.dev)
            self.n_edges = int(self.g.rowptr[self.U].item())
        for _ in range(n_steps):
            ops.sample_bpr(self.g, self.U, self.I, self.n_edges, seed, self.sampler_state,
                           self.b_users, self.b_pos, self.b_neg)
            if use_graph:
                if self._graph is None:
                    self.capture()
                self._graph.replay()
            else:
                self._step_body()
        return self.loss

    # ---- checkpoints in the reference's format (main.py:550 saves model.state_dict()) ---------
    def _full_table(self):
        return self.P

    def _local_columns(self, full):
        return full

    def _full_content(self):
        return self.fusion["C"]

    def state_dict(self):
        """The parameters under the reference's ``state_dict`` keys (``models/lightgcn.py:15-17``,
        ``models/lightgcn_fusion.py:20-35``; same key order), as CPU tensors -- what
        ``torch.save(model.state_dict(), ...)`` of reference ``main.py:550`` would hold, whatever
        the sharding.  Collective when the engine is sharded (every rank gets the full dict)."""
        from collections import OrderedDict
        U, I = self.U, self.I
        P = self._full_table().detach().cpu()
        sd = OrderedDict()
        if self.fusion is None:
            sd["user_embedding.weight"] = P[:U].clone()
            sd["brand_embedding.weight"] = P[U + I:].clone()
            sd["item_embedding.weight"] = P[U:U + I].clone()
        else:
            sd["item_content_embedding"] = self._full_content().detach().cpu().clone()
            sd["user_embedding.weight"] = P[:U].clone()
            sd["item_id_embedding.weight"] = P[U:U + I].clone()
            sd["brand_embedding.weight"] = P[U + I:].clone()
            sd["item_fusion_layer.weight"] = self.fusion["W"].detach().cpu().clone()
            sd["item_fusion_layer.bias"] = self.fusion["b"].detach().cpu().clone()
        return sd

    def load_state_dict(self, sd):
        """Load a reference-format checkpoint (``main.py:571``) into the (possibly sharded)
        tables.  Optimizer state is not part of the reference's checkpoints and is left as is."""
        U, I, B = self.U, self.I, self.B
        item_key = "item_embedding.weight" if self.fusion is None else "item_id_embedding.weight"
        parts = [sd["user_embedding.weight"], sd[item_key], sd["brand_embedding.weight"]]
        for t, rows in zip(parts, (U, I, B)):
            if tuple(t.shape[:1]) != (rows,):
                raise LgcnError(f"checkpoint table has {t.shape[0]} rows, expected {rows}")
        full = torch.cat([t.to(torch.float32) for t in parts], 0)
        self.P.copy_(self._local_columns(full))
        if self.fusion is not None:
            self.fusion["W"].copy_(sd["item_fusion_layer.weight"])
            self.fusion["b"].copy_(sd["item_fusion_layer.bias"])

    def bpr_loss(self, users, pos, neg):
        """Loss only (no gradient, no update) on the current parameters."""
        F = self.propagate()
        return ops.bpr_fused(F, self.P, users, pos, neg, self.U, self.lam)

    def _rate(self, F, users, mask_rowptr, mask_col, k, batch_users=None):
        """Batched full-rank rating against the [N, d] table ``F`` with a rater (prepared bf16 item
        tiles + workspace) that lives across calls; ``propagate`` invalidates the prepared tiles."""
        if (users.numel() and (self._idx_checked_eval is not users)) or _CHECK_ALWAYS:
            self.idx_status.zero_()
            ops.check_indices(self.idx_status, (users, 0, self.U))
            if int(self.idx_status.item()) != 0:
                raise IndexError("evaluation user index out of range")
            self._idx_checked_eval = users
        Fu, Fi = F[:self.U], F[self.U:self.U + self.I]
        d = F.shape[1]
        rater = None
        if d in (64, 128) and self.I >= ops.TC_MIN_ITEMS and k <= 32 and users.numel() > 0:
            want = min(users.numel(), batch_users or ops.TC_WAVE_USERS * ops.TC_BATCH_WAVES)
            if self._tc is None or self._tc.max_users < want or self._tc.d != d:
                self._tc = ops.TcRater(self.I, d, self.dev, want)
            rater = self._tc
        return ops.score_topk(Fu, Fi, users, mask_rowptr, mask_col, k, rater=rater,
                              batch_users=batch_users)

    def rate_topk(self, users, mask_rowptr=None, mask_col=None, k=20, propagate=True, batch_users=None):
        """Full-rank top-k item ids for ``users`` with their train items excluded (reference
        ``main.py:413-426``); any number of users -- the sweep runs in user batches."""
        F = self.propagate() if propagate else self.F
        return self._rate(F, users, mask_rowptr, mask_col, k, batch_users)

    def evaluate(self, eval_users, targets, mask_rowptr, mask_col, k=20, propagate=True, batch_users=None):
        """recall@k / NDCG@k as reference ``main.py:404-439``.  Returns (recall, ndcg, ids)."""
        if propagate:
            self.propagate()
        nu = eval_users.numel()
        sums = torch.zeros(2, dtype=torch.float64, device=self.dev)
        ids, _ = self.rate_topk(eval_users, mask_rowptr, mask_col, k, propagate=False,
                                batch_users=batch_users)
        ops.eval_metrics(ids, targets, sums)
        s = sums.cpu().numpy()
        return float(s[0] / nu), float(s[1] / nu), ids
