"""Profiling driver: the ADAM-epilogue SpMM (last backward hop) at a workload's shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_recommendation_b200 import ops, synth  # noqa: E402
from gcn_recommendation_b200.graph import NormAdjCSR  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "amazon"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
U, I, B, total, d, K = synth.SHAPES[workload]
if len(sys.argv) > 3:
    d = int(sys.argv[3])
inter = synth.generate_device(workload, dev, seed=0)
tu, ti, _, _ = synth.split_validation_device(inter)
g = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev)
del inter, tu, ti
N = U + I + B
x = torch.randn((N, d), device=dev) * 1e-3
p = torch.randn((N, d), device=dev)
m = torch.zeros((N, d), device=dev)
v = torch.full((N, d), 1e-6, device=dev)
add = torch.zeros((N, d), device=dev)
rows = torch.randint(0, N, (6144,), device=dev)
add[rows] = torch.randn((6144, d), device=dev)
flag = torch.zeros(N + 32, dtype=torch.uint8, device=dev)
flag[rows] = 1
zero_row = torch.zeros(256, device=dev)
sc = torch.tensor([1e-3, 1.0], device=dev)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
ev[0].record()
for i in range(n):
    ops.spmm_adam(g, x, p, m, v, sc, addend=add, addend_rowflag=flag, zero_row=zero_row)
    ev[i + 1].record()
torch.cuda.synchronize()
print(f"{workload} adam d={d} ms={[round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(n)]}")
