"""Profiling driver: full evaluate() (propagation + tensor-core rating + metrics) on a synthetic graph."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_recommendation_b200 import ops, synth  # noqa: E402
from gcn_recommendation_b200.engine import LightGCNEngine, build_mask_csr, xavier_uniform_table  # noqa: E402
from gcn_recommendation_b200.graph import NormAdjCSR  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "amazon"
nu = int(sys.argv[2]) if len(sys.argv) > 2 else 18944
dev = torch.device("cuda:0")
U, I, B, total, d, K = synth.SHAPES[workload]
inter = synth.generate_device(workload, dev, seed=0)
tu, ti, vu, vi = synth.split_validation_device(inter)
g = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev)
eng = LightGCNEngine(g, U, I, B, K, xavier_uniform_table([U, I, B], d, dev))
eu, tg = vu[:nu].contiguous(), vi[:nu].contiguous()
mr, mc = build_mask_csr(eu.cpu().numpy(), tu.cpu().numpy(), ti.cpu().numpy(), U, dev)
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ops.PROFILE = []
    eng.propagate()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ids, _ = eng.rate_topk(eu, mr, mc, 20, propagate=False)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    parts = {t: round(a.elapsed_time(b), 2) for t, a, b in ops.PROFILE if t.startswith("score")}
    ops.PROFILE = None
    print(f"{workload} users={nu}: propagate {1e3 * (t1 - t0):.1f} ms, rate_topk {1e3 * (t2 - t1):.1f} ms {parts} "
          f"-> {nu / (t2 - t0):.0f} users/s incl. propagation, {nu / (t2 - t1):.0f} users/s rating only; {ops.STATS}")
