"""A/B sweep of the column residency classes (include/lgcn.h LGCN_COL_*): SpMM launch time at the
Amazon shape for several L2 budgets of HOT columns, with and without evict_first for the
unclassified columns.  One process, the graph is built once.

    python profiles/prof_hot.py [workload] [d,d,...] [mode,mode,...] [mb,mb,...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_recommendation_b200 import _lib, ops, synth  # noqa: E402
from gcn_recommendation_b200.graph import NormAdjCSR  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "amazon"
dims = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "128").split(",")]
modes = (sys.argv[3] if len(sys.argv) > 3 else "plain").split(",")
budgets = [float(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "0,32,64,96").split(",")]
n = 6
dev = torch.device("cuda:0")
U, I, B, total, _, K = synth.SHAPES[workload]
inter = synth.generate_device(workload, dev, seed=0)
tu, ti, _, _ = synth.split_validation_device(inter)
g = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev)
del inter, tu, ti
N = U + I + B
for d in dims:
    x = torch.randn((N, d), device=dev)
    y = torch.empty_like(x)
    add = torch.randn((N, d), device=dev)
    p, m, v = torch.randn((N, d), device=dev), torch.zeros((N, d), device=dev), torch.ones((N, d), device=dev)
    sc = torch.tensor([1e-3, 1.0], device=dev)
    for mode in modes:
        for mb in budgets:
            for cold_first in (0, 1):
                if mb == 0 and cold_first:
                    continue
                ops.HOT_BYTES = int(mb * (1 << 20))
                ops.SPMM_FLAGS_EXTRA = _lib.SPMM_F_COLD_FIRST if cold_first else 0
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
                torch.cuda.synchronize()
                ev[0].record()
                for i in range(n):
                    if mode == "plain":
                        ops.spmm(g, x, out=y)
                    elif mode == "add":
                        ops.spmm(g, x, out=y, addend=add)
                    elif mode == "mean":
                        ops.spmm(g, x, out=y, mean_layers=[add, x, add, x][:K])
                    elif mode == "adam":
                        ops.spmm_adam(g, x, p, m, v, sc, addend=add)
                    ev[i + 1].record()
                torch.cuda.synchronize()
                ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(1, n))
                print(f"{workload} d={d} {mode:5s} hot_mb={mb:5.0f} n_hot={getattr(g, 'n_hot', 0):7d} "
                      f"cold_first={cold_first} median_ms={ms[len(ms) // 2]:.3f} min_ms={ms[0]:.3f}", flush=True)
    del x, y, add, p, m, v
