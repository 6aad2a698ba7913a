"""Profiling driver: tensor-core rating sweep on random tables (run under ncu via gpurun).

    python profiles/prof_score.py [n_users] [n_items] [d]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_recommendation_b200 import ops  # noqa: E402

nu = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
ni = int(sys.argv[2]) if len(sys.argv) > 2 else 4_400_000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 128
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
Fu = torch.randn((nu, d), device=dev, generator=g) * 0.1
Fi = torch.randn((ni, d), device=dev, generator=g) * 0.1 * (0.2 + 2 * torch.rand((ni, 1), device=dev, generator=g))
users = torch.arange(nu, device=dev)
for it in range(3):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.PROFILE = []
    ev0.record()
    ids, sc = ops.score_topk(Fu, Fi, users, None, None, 20, tensor_cores=True)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    parts = {t: round(a.elapsed_time(b), 2) for t, a, b in ops.PROFILE}
    ops.PROFILE = None
    print(f"users={nu} items={ni} d={d}: {ms:.2f} ms  {nu / ms * 1e3:.0f} users/s  "
          f"{2.0 * nu * ni * d / ms / 1e9:.1f} TFLOP/s  parts={parts} "
          f"filter TFLOP/s={2.0 * nu * ni * d / parts['score_tc_filter_refine'] / 1e9:.1f} stats={ops.STATS}")
