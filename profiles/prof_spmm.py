"""Profiling driver: a few SpMM launches on a synthetic graph (run under ncu via gpurun).

    python profiles/prof_spmm.py [workload] [mode] [launches]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_recommendation_b200 import ops, synth  # noqa: E402
from gcn_recommendation_b200.graph import NormAdjCSR  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "amazon"
mode = sys.argv[2] if len(sys.argv) > 2 else "plain"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda:0")
if os.environ.get("LGCN_L2_FETCH"):
    # experiment: device-wide L2 fetch granularity (cudaLimitMaxL2FetchGranularity
    # = 0x05; 32 / 64 / 128 bytes) -- narrow (64-byte) rows pay a 128-byte DRAM fetch by default
    import ctypes
    torch.zeros(1, device=dev)
    rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["LGCN_L2_FETCH"])))
    val = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
    print(f"cudaLimitMaxL2FetchGranularity <- {os.environ['LGCN_L2_FETCH']}: rc={rc}, now {val.value}")
# LGCN_SPMM_FLAGS (read by ops): 2 = chunk kernel instead of ring, 16 = ring for the ADAM epilogue too
U, I, B, total, d, K = synth.SHAPES[workload]
if len(sys.argv) > 4:
    d = int(sys.argv[4])
inter = synth.generate_device(workload, dev, seed=0)
tu, ti, _, _ = synth.split_validation_device(inter)
if os.environ.get("LGCN_RELABEL"):
    # experiment: relabel the USERS so that users sharing an item are neighbours
    # in the row order -> the item row they gather is re-used while it is still in L2.  Keys:
    #   coldest      : the user's lowest-degree item
    #   hot:<n>      : the user's highest-degree item outside the <n> hottest (those stay L2-resident)
    rl = os.environ["LGCN_RELABEL"]
    deg_i = torch.bincount(ti, minlength=I)
    if rl == "coldest":
        score = deg_i[ti] * I + ti
    else:
        nres = int(rl.split(":")[1])
        rank = torch.empty(I, dtype=torch.int64, device=dev)
        rank[torch.argsort(deg_i, descending=True, stable=True)] = torch.arange(I, device=dev)
        r = rank[ti]
        score = torch.where(r >= nres, r, r + 10 * I)
    key = torch.full((U,), 1 << 62, dtype=torch.int64, device=dev).scatter_reduce_(0, tu, score, "amin")
    order = torch.argsort(key, stable=True)
    newid = torch.empty(U, dtype=torch.int64, device=dev)
    newid[order] = torch.arange(U, device=dev)
    tu = newid[tu]
    print(f"relabelled users by {rl}")
    del deg_i, score, key, order, newid
g = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev)
del inter, tu, ti
N = U + I + B
x = torch.randn((N, d), device=dev)
y = torch.empty_like(x)
add = torch.randn((N, d), device=dev) if mode != "plain" else None
if mode == "addflag":     # a later Horner hop: dense x, addend g' with 3*2048 non-zero rows, flagged
    rows = torch.randint(0, N, (6144,), device=dev)
    add.zero_()
    add[rows] = torch.randn((6144, d), device=dev)
    flag = torch.zeros(N + 32, dtype=torch.uint8, device=dev)
    flag[rows] = 1
    zero_row = torch.zeros(256, device=dev)
if mode == "hop2":        # second Horner hop: x = first hop's output, ~20 % non-zero rows, flagged; dense out
    live = torch.rand(N, device=dev) < 0.2
    x[~live] = 0
    flag = torch.zeros(N + 32, dtype=torch.uint8, device=dev)
    flag[:N] = live.to(torch.uint8)
    rows = torch.randint(0, N, (6144,), device=dev)
    add = torch.zeros((N, d), device=dev)
    add[rows] = torch.randn((6144, d), device=dev)
    aflag = torch.zeros(N + 32, dtype=torch.uint8, device=dev)
    aflag[rows] = 1
    zero_row = torch.zeros(256, device=dev)
if mode in ("hop1", "hop1s"):        # the first Horner hop: x = g' has 3*2048 non-zero rows, flagged
    rows = torch.randint(0, N, (6144,), device=dev)
    x.zero_()
    x[rows] = torch.randn((6144, d), device=dev)
    flag = torch.zeros(N + 32, dtype=torch.uint8, device=dev)
    flag[rows] = 1
    yflag = torch.zeros(N + 32, dtype=torch.uint8, device=dev)
    zero_row = torch.zeros(256, device=dev)
    add = x
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
ev[0].record()
for i in range(n):
    if mode == "plain":
        ops.spmm(g, x, out=y)
    elif mode == "add":
        ops.spmm(g, x, out=y, addend=add)
    elif mode == "addflag":
        ops.spmm(g, x, out=y, addend=add, addend_rowflag=flag, zero_row=zero_row)
    elif mode == "hop1":
        ops.spmm(g, x, out=y, addend=add, x_rowflag=flag, addend_rowflag=flag, zero_row=zero_row)
    elif mode == "hop2":
        ops.spmm(g, x, out=y, addend=add, x_rowflag=flag, addend_rowflag=aflag, zero_row=zero_row)
    elif mode == "hop1s":     # + sparse output (all-zero rows not written, reported in yflag)
        ops.spmm(g, x, out=y, addend=add, x_rowflag=flag, addend_rowflag=flag, zero_row=zero_row,
                 y_rowflag=yflag)
    elif mode == "mean":
        ops.spmm(g, x, out=y, mean_layers=[add, x, add, x][:K])
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
print(f"{workload} {mode} d={d} N={N} nnz={g.nnz} n_long={g.n_long} n_seg={g.n_seg} ms={ms}")
