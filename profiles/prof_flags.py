"""Sparsity of the backward hops at a workload's shape: rows flagged by the BPR scatter (g'), rows
the first Horner hop can make non-zero, and the share of gathers that hit them."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gcn_recommendation_b200 import synth  # noqa: E402
from gcn_recommendation_b200.engine import LightGCNEngine, xavier_uniform_table  # noqa: E402
from gcn_recommendation_b200.graph import NormAdjCSR  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "amazon"
dev = torch.device("cuda:0")
U, I, B, total, d, K = synth.SHAPES[workload]
inter = synth.generate_device(workload, dev, seed=0)
tu, ti, _, _ = synth.split_validation_device(inter)
g = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev)
table = xavier_uniform_table([U, I, B], d, dev, torch.Generator(device=dev).manual_seed(42))
eng = LightGCNEngine(g, U, I, B, K, table)
batches = bench.make_batches(tu.cpu(), ti.cpu(), I, 3, 1, device=dev)
N = U + I + B
for u, p, n in batches:
    eng.b_users.copy_(u); eng.b_pos.copy_(p); eng.b_neg.copy_(n)
    F = eng.propagate()
    eng._bpr(F, True)
    rf = eng.rowflag[:N].bool()
    deg = (g.rowptr[1:] - g.rowptr[:-1]).long()
    nbr = torch.zeros(N, dtype=torch.bool, device=dev)
    rows = torch.repeat_interleave(torch.arange(N, device=dev), deg)
    live_e = rf[g.col.long()]
    nbr[rows[live_e]] = True
    f2 = nbr | rf
    live2 = f2[g.col.long()]
    print(f"{workload}: g' rows {int(rf.sum())}  hop-1 live gathers {float(live_e.float().mean()):.4f}  "
          f"hop-1 non-zero output rows {int(f2.sum())} ({float(f2.float().mean()):.3f}; users "
          f"{float(f2[:U].float().mean()):.3f}, items {float(f2[U:U+I].float().mean()):.3f})  "
          f"hop-2 live gathers {float(live2.float().mean()):.3f}", flush=True)
    from gcn_recommendation_b200 import ops
    ops.zero_rows(eng.G1, eng.G2, u, p, n, U, rowflag=eng.rowflag)
