"""Multi-GPU LightGCN: one process per GPU, ``torch.distributed`` (NCCL over NVLink) plumbing.

Two partitionings of the same step (DESIGN.md "Multi-GPU"):

* :class:`FeatureShardedEngine` -- every rank owns ``d/P`` feature COLUMNS of every table row
  (parameters, Adam moments, layers, gradients) and the whole CSR.  The propagation
  ``A_hat @ X`` is column-separable, so forward, backward and Adam need NO communication at
  all; the only exchange is one all-reduce of ``3*batch`` partial dot products per step (the
  BPR scores need the full feature dimension).  Evaluation all-gathers the final table once
  and shards the users.
* :class:`RowShardedEngine` -- the north-star layout: every rank owns a block of ROWS (of the
  graph and of every table) and the layer input is all-gathered before each SpMM.  At the
  Amazon shape this moves ``T*(P-1)/P`` = 6.6 GB per rank per layer over NVLink against ~1 ms
  of local SpMM, i.e. it is communication bound; it is kept as the measured comparison.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from ._lib import LgcnError
from .engine import LightGCNEngine


def column_shard(table, rank, world):
    """This rank's ``d/world`` columns of a [N,d] table as a contiguous [N, d/world] tensor."""
    d = table.shape[1]
    if d % world or (d // world) not in (16, 32, 64, 128, 256):
        raise LgcnError(f"d={d} cannot be split over {world} ranks (local width must be 16..256)")
    dl = d // world
    return table[:, rank * dl:(rank + 1) * dl].contiguous()


def cols_to_rows(cols, a2a_in, a2a_out, rows_out, alltoall):
    """[I, d/P] column shard of an item table -> full rows [ipr, d] of this rank's item block
    (ipr = ceil(I/P) items per rank).  ``a2a_in`` / ``a2a_out``: [P, ipr, d/P] exchange buffers;
    block s of the input goes to rank s, block s of the output came from rank s."""
    world, ipr, dl = a2a_in.shape
    a2a_in.view(world * ipr, dl)[:cols.shape[0]].copy_(cols)
    alltoall(a2a_out, a2a_in)
    rows_out.view(ipr, world, dl).copy_(a2a_out.permute(1, 0, 2))
    return rows_out


def rows_to_cols(rows, a2a_in, a2a_out, cols_out, alltoall):
    """Inverse of :func:`cols_to_rows`: full rows of this rank's item block -> column shard."""
    world, ipr, dl = a2a_in.shape
    a2a_in.copy_(rows.view(ipr, world, dl).permute(1, 0, 2))
    alltoall(a2a_out, a2a_in)
    cols_out.copy_(a2a_out.view(world * ipr, dl)[:cols_out.shape[0]])
    return cols_out


class FeatureShardedEngine(LightGCNEngine):
    """LightGCN step with the feature dimension split over the ranks of ``group``."""

    def __init__(self, g, num_users, num_items, num_brands, n_layers, table_local, group=None,
                 allreduce=None, world=None, rank=None, **kw):
        self.group = group
        self.world = world if world is not None else dist.get_world_size(group)
        self.rank = rank if rank is not None else dist.get_rank(group)
        self._allreduce = allreduce or (lambda t: dist.all_reduce(t, group=self.group))
        super().__init__(g, num_users, num_items, num_brands, n_layers, table_local, **kw)
        self.launches_per_step += 1

    def _alloc_batch(self, bs):
        super()._alloc_batch(bs)
        self.dots = torch.empty(3 * bs, dtype=torch.float32, device=self.dev)

    def _extra_batch_buffers(self):
        return ("dots",)

    # ---- LightGCN_Fusion item block (reference models/lightgcn_fusion.py:45-49) -------------
    # The projection mixes ALL d features of an item row, so it cannot run on column shards.
    # It is sharded by ITEM instead: rank r owns the content rows of item block
    # r.  One all-to-all turns the column shards of the id embeddings into full rows of the
    # rank's item block, the tcgen05 projection kernel runs on them with the replicated W / b,
    # and a second all-to-all scatters the projected rows back into column shards.  The backward
    # pass mirrors it; dW / db are summed with one small all-reduce, so W and b (and their Adam
    # moments) stay replicated bit-identically.  Volume: I*d*4*(P-1)/P^2 bytes per rank and
    # exchange (246 MB at the Amazon shape on 8 GPUs), 4 exchanges per step.
    def _init_fusion(self, f):
        C, W, b = f["content"], f["weight"], f["bias"]
        P_ = self.world
        self.d_full = self.d * P_
        self.ipr = -(-self.I // P_)                                # items per rank (padded)
        self.i0 = min(self.I, self.rank * self.ipr)
        self.i_loc = min(self.I, self.i0 + self.ipr) - self.i0
        if (C.shape[0] != self.i_loc or W.shape != (self.d_full, self.d_full + C.shape[1])
                or b.shape != (self.d_full,)):
            raise LgcnError("fusion: content must hold this rank's item block "
                            f"[{self.i_loc}, c], weight [d, d+c] and bias [d] with d={self.d_full}")
        z = torch.zeros_like
        new = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=self.dev)  # noqa: E731
        self.fusion = dict(C=C.contiguous(), W=W, b=b, mW=z(W), vW=z(W), mb=z(b), vb=z(b),
                           gW=z(W), gb=z(b), g_eid=new(self.I, self.d))
        self.H, self.gH = new(self.I, self.d), new(self.I, self.d)     # column shards of H and dL/dH
        self.a2a_in, self.a2a_out = new(P_, self.ipr, self.d), new(P_, self.ipr, self.d)
        self.E_rows, self.H_rows = new(self.ipr, self.d_full), new(self.ipr, self.d_full)
        self.G_rows, self.GE_rows = new(self.ipr, self.d_full), new(self.ipr, self.d_full)
        self._alltoall = f.get("alltoall") or (
            lambda out, inp: dist.all_to_all_single(out, inp, group=self.group))
        self._allreduce_sum = f.get("allreduce") or self._allreduce

    def _count_launches(self):
        n = super()._count_launches()
        return n + (1 if self.fusion is not None else 0)           # + exchanges are NCCL's

    def _cols_to_rows(self, cols, rows_out):
        return cols_to_rows(cols, self.a2a_in, self.a2a_out, rows_out, self._alltoall)

    def _rows_to_cols(self, rows, cols_out):
        return rows_to_cols(rows, self.a2a_in, self.a2a_out, cols_out, self._alltoall)

    def layer0(self):
        if self.fusion is None:
            return self.P, None
        U, I, f, n = self.U, self.I, self.fusion, self.i_loc
        self._cols_to_rows(self.P[U:U + I], self.E_rows)
        if n > 0:
            ops.fusion_proj_fwd(self.E_rows[:n], f["C"], f["W"], f["b"], out=self.H_rows[:n])
        self._rows_to_cols(self.H_rows, self.H)
        return self.P, (self.H, U)

    def _fusion_backward(self, gH):
        U, I, f, n = self.U, self.I, self.fusion, self.i_loc
        f["gW"].zero_()
        f["gb"].zero_()
        self._cols_to_rows(gH, self.G_rows)                          # dL/dH rows of the item block
        if n > 0:                                                    # E_rows / H_rows: kept from layer0
            ops.fusion_proj_bwd(self.E_rows[:n], f["C"], f["W"], self.H_rows[:n], self.G_rows[:n],
                                g_eid=self.GE_rows[:n], gW=f["gW"], gb=f["gb"])
        self._allreduce_sum(f["gW"])
        self._allreduce_sum(f["gb"])
        self._rows_to_cols(self.GE_rows, f["g_eid"])

    def _bpr_term(self, F, pos, neg, item_offset, lam, scale, gp_includes_gf, loss_out):
        u = self.b_users
        ops.bpr_partial(F, self.P, u, pos, neg, item_offset, self.dots)
        self._allreduce(self.dots)                     # the only collective of the term
        ops.bpr_apply(F, self.P, u, pos, neg, item_offset, lam, self.dots, grad_scale=scale,
                      gF=self.G1, gP=self.G2, gp_includes_gf=gp_includes_gf, sample_ws=self.sample_ws,
                      loss_out=loss_out, rowflag=self.rowflag)

    # ---- reference-format checkpoints from column shards -------------------------
    def _gather_columns(self, local):
        rows, dl = local.shape
        out = torch.empty((self.world * rows, dl), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        return out.view(self.world, rows, dl).permute(1, 0, 2).reshape(rows, self.world * dl).contiguous()

    def _full_table(self):
        return self._gather_columns(self.P)

    def _local_columns(self, full):
        dl = self.d
        if full.shape[1] != dl * self.world:
            raise LgcnError(f"checkpoint width {full.shape[1]} != {dl * self.world}")
        return full[:, self.rank * dl:(self.rank + 1) * dl]

    def _full_content(self):
        C = self.fusion["C"]                                        # this rank's item block
        pad = torch.zeros((self.ipr, C.shape[1]), dtype=C.dtype, device=C.device)
        pad[:C.shape[0]].copy_(C)
        out = torch.empty((self.world * self.ipr, C.shape[1]), dtype=C.dtype, device=C.device)
        dist.all_gather_into_tensor(out, pad, group=self.group)
        return out[:self.I]

    def gather_final_table(self):
        """All-gather the propagated table over the feature dimension -> [N, d] on every rank."""
        F = self.propagate()
        return self._gather_columns(F)

    def evaluate(self, eval_users, targets, mask_rowptr, mask_col, k=20, propagate=True, batch_users=None):
        """User-sharded full-rank evaluation (reference ``main.py:404-439``): the caller passes
        THIS rank's users / targets / mask rows; hit and DCG sums are all-reduced."""
        if propagate or getattr(self, "_F_full", None) is None:
            self._F_full = None                     # release the old gathered table first
            self._F_full = self.gather_final_table()
            if self._tc is not None:
                self._tc.prepared_for = None
        F = self._F_full
        ids, _ = self._rate(F, eval_users, mask_rowptr, mask_col, k, batch_users)
        sums = torch.zeros(3, dtype=torch.float64, device=self.dev)
        ops.eval_metrics(ids, targets, sums[:2])
        sums[2] = eval_users.numel()
        dist.all_reduce(sums, group=self.group)
        s = sums.cpu().numpy()
        return float(s[0] / s[2]), float(s[1] / s[2]), ids


class RowShardedEngine:
    """Row-sharded propagation with a per-layer NCCL all-gather (BASELINE.json north_star).

    Rank r owns rows ``[r*rows_per_rank, (r+1)*rows_per_rank)`` of the node space padded to a
    multiple of the world size: its CSR rows, parameters, Adam moments and layer buffers.  Before
    every SpMM the layer input is all-gathered into a full [N_pad, d] buffer.  Each rank forms the
    whole batch's BPR terms redundantly from the gathered table and keeps only the gradient rows
    it owns, so the loss needs no collective.
    """

    def __init__(self, g_full, num_users, num_items, num_brands, n_layers, table_full, group=None,
                 lr=1e-3, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, batch_size=2048):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.U, self.I, self.B = int(num_users), int(num_items), int(num_brands)
        self.N = self.U + self.I + self.B
        self.K = int(n_layers)
        self.d = int(table_full.shape[1])
        self.dev = table_full.device
        self.lr, self.lam, self.betas, self.eps = float(lr), float(weight_decay), betas, float(eps)
        P_ = self.world
        self.rpr = -(-self.N // P_)                               # rows per rank (padded)
        self.r0 = self.rank * self.rpr
        self.r1 = min(self.N, self.r0 + self.rpr)
        self.nloc = max(0, self.r1 - self.r0)
        self.g = g_full.row_shard(self.r0, self.r1) if self.nloc > 0 else None
        z = lambda: torch.zeros((self.rpr, self.d), dtype=torch.float32, device=self.dev)  # noqa: E731
        self.P = z()
        self.P[:self.nloc].copy_(table_full[self.r0:self.r1])
        self.m, self.v = z(), z()
        self.layers = [z() for _ in range(max(self.K - 1, 1))]     # local E_1..E_{K-1}
        self.Floc = z()
        self.G1, self.G2 = z(), z()
        self.acc = [z(), z()]
        full = lambda: torch.empty((self.rpr * P_, self.d), dtype=torch.float32, device=self.dev)  # noqa: E731
        self.Xa, self.Xb = full(), full()                          # gathered tables (ping-pong)
        self.P_full = full()
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.adam_scalars = torch.zeros(2, dtype=torch.float32, device=self.dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self.bs = int(batch_size)
        self.sample_ws = torch.empty(2 * self.bs, dtype=torch.float32, device=self.dev)
        self.idx_status = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._checked_bs = -1
        self.Gfull = torch.zeros((self.rpr * P_, self.d), dtype=torch.float32, device=self.dev)
        self.G2full = torch.zeros((self.rpr * P_, self.d), dtype=torch.float32, device=self.dev)
        per_spmm = ops.spmm_launches(self.g, self.d) if self.g is not None else 1
        self.launches_per_step = 2 * self.K * per_spmm + 2 + 1 + 1
        self.collectives_per_step = 2 * self.K

    def _gather(self, local, out):
        dist.all_gather_into_tensor(out, local, group=self.group)
        return out

    def propagate(self):
        K = self.K
        x = self._gather(self.P, self.P_full)                      # layer 0 on every rank
        mean_layers = [self.P]
        for k in range(K - 1):
            ops.spmm(self.g, x, out=self.layers[k])
            mean_layers.append(self.layers[k])
            x = self._gather(self.layers[k], self.Xa if k % 2 == 0 else self.Xb)
        ops.spmm(self.g, x, out=self.Floc, mean_layers=mean_layers)
        return self.Floc

    def bpr_step(self, users, pos, neg, use_graph=False):
        K, U = self.K, self.U
        dev = self.dev
        bs = users.numel()
        if not (pos.numel() == bs and neg.numel() == bs):
            raise LgcnError("users, pos and neg must have the same length")
        if bs != self.bs:                          # the fused BPR kernel writes sample_ws[0 : 2*bs]
            self.bs = bs
            self.sample_ws = torch.empty(2 * bs, dtype=torch.float32, device=dev)
            self._checked_bs = -1
        u, p, n = users.to(dev, non_blocking=True), pos.to(dev, non_blocking=True), neg.to(dev, non_blocking=True)
        if self._checked_bs != bs:
            self.idx_status.zero_()
            ops.check_indices(self.idx_status, (u, 0, self.U), (p, 0, self.I), (n, 0, self.I))
            if int(self.idx_status.item()) != 0:
                raise IndexError("batch indices are out of range")
            self._checked_bs = bs
        self._F_full = None                        # it is a view of a ping-pong buffer
        self.propagate()
        F_full = self._gather(self.Floc, self.Xa if (K - 1) % 2 == 0 else self.Xb)
        # every rank forms the whole batch redundantly; gradients land in full-size scratch
        # tables of which only the owned row block is used afterwards
        ops.bpr_fused(F_full, self.P_full, u, p, n, U, self.lam, grad_scale=1.0 / (K + 1),
                      gF=self.Gfull, gP=self.G2full, gp_includes_gf=True, sample_ws=self.sample_ws,
                      loss_out=self.loss)
        ops.adam_tick(self.step_dev, self.adam_scalars, self.lr, self.betas)
        g1 = self.Gfull[self.r0:self.r0 + self.rpr]
        g2 = self.G2full[self.r0:self.r0 + self.rpr]
        x = self.Gfull                                             # acc_0 = g' is already global
        for k in range(K - 1):
            ops.spmm(self.g, x, out=self.acc[k % 2], addend=g1)
            x = self._gather(self.acc[k % 2], self.Xa if k % 2 == 0 else self.Xb)
        ops.spmm_adam(self.g, x, self.P, self.m, self.v, self.adam_scalars, addend=g2,
                      betas=self.betas, eps=self.eps)
        ops.zero_rows(self.Gfull, self.G2full, u, p, n, U)
        return self.loss

    # ---- reference-format checkpoints and evaluation from row blocks ----------------------------
    fusion = None

    def _full_table(self):
        """All-gather the parameter row blocks -> [N, d] on every rank."""
        return self._gather(self.P, self.P_full)[:self.N]

    def _local_columns(self, full):
        """The padded row block of this rank out of a full [N, d] table (name kept from the
        feature-sharded engine: it is what ``load_state_dict`` copies into ``P``)."""
        blk = torch.zeros((self.rpr, self.d), dtype=torch.float32, device=self.dev)
        if self.nloc > 0:
            blk[:self.nloc].copy_(full[self.r0:self.r1])
        return blk

    state_dict = LightGCNEngine.state_dict
    load_state_dict = LightGCNEngine.load_state_dict

    def gather_final_table(self):
        self.propagate()
        return self._gather(self.Floc, self.Xa)[:self.N]

    def evaluate(self, eval_users, targets, mask_rowptr, mask_col, k=20, propagate=True, batch_users=None):
        """User-sharded full-rank evaluation (reference ``main.py:404-439``): this rank's users /
        targets / mask rows against the all-gathered final table; sums all-reduced."""
        if propagate or getattr(self, "_F_full", None) is None:
            self._F_full = self.gather_final_table()
        F = self._F_full
        ids, _ = ops.score_topk(F[:self.U], F[self.U:self.U + self.I], eval_users, mask_rowptr,
                                mask_col, k, batch_users=batch_users)
        sums = torch.zeros(3, dtype=torch.float64, device=self.dev)
        ops.eval_metrics(ids, targets, sums[:2])
        sums[2] = eval_users.numel()
        dist.all_reduce(sums, group=self.group)
        s = sums.cpu().numpy()
        return float(s[0] / s[2]), float(s[1] / s[2]), ids
