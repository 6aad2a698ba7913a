"""LRU model of the L2 behind one Amazon-shape lgcn_spmm launch (CPU, numpy; no GPU needed).

    python profiles/micro/lru_model.py [amazon_16th|amazon]

The kernel walks the rows in order; every entry gathers one 512-byte row.  An access HITS iff the
traffic that went through the L2 since the previous access to the same row (gathers + the streamed
output / entry bytes that ride along) is below the capacity.  The capacity is scaled with the graph
(1/16-scale graph -> 1/16 of the L2), Zipf(0.8) popularity shares are scale free.  Output: hit
fraction of the user-side gathers (item rows, popularity skewed) and of the item-side gathers (user
rows, each gathered ~2.9 times at random distances), for the original node order and for users
relabelled so that users sharing an item are neighbours.  Measured counterparts:
profiles/r02_zipf_l2_microbench.txt (ncu: 28.7 % / 19 % user-side hits without / with the store
stream; relabelling: -3.7 % of the plain launch).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gcn_recommendation_b200 import synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "amazon_16th"
scale = {"amazon_16th": 16, "amazon_64th": 64, "amazon": 1}[shape]
U, I, B, total, d, K = synth.SHAPES[shape]
inter = synth.generate(shape, seed=0)
tu, ti, _, _ = inter.split_validation()
ROW = 512


def hits(seq, extra_per_access, cap_bytes):
    n = len(seq)
    order = np.argsort(seq, kind="stable")
    s = seq[order]
    prev = np.full(n, -1, np.int64)
    same = s[1:] == s[:-1]
    prev[order[1:][same]] = order[:-1][same]
    dist = np.where(prev >= 0, np.arange(n) - prev, 1 << 60)
    return float((dist * (ROW + extra_per_access) < cap_bytes).mean())


def evaluate(tu, ti, label, caps=(48, 64, 96)):
    seq_u = ti[np.lexsort((ti, tu))]          # user rows gather item rows
    seq_i = tu[np.lexsort((tu, ti))]          # item rows gather user rows
    cells = []
    for cap in caps:
        c = cap * 2 ** 20 / scale
        hu = hits(seq_u, ROW * U / len(tu) + 8, c)
        hi = hits(seq_i, ROW * I / len(tu) + 8, c)
        cells.append(f"[{cap} MB: user-side {hu:.3f} item-side {hi:.3f}]")
    print(f"{label:46s}", " ".join(cells), flush=True)


def relabel(key):
    order = np.argsort(key, kind="stable")
    new = np.empty(U, np.int64)
    new[order] = np.arange(U)
    return new


evaluate(tu, ti, "original order")
deg_i = np.bincount(ti, minlength=I)
rank_i = np.empty(I, np.int64)
rank_i[np.argsort(-deg_i, kind="stable")] = np.arange(I)             # 0 = hottest item
first = lambda su: np.concatenate([[True], su[1:] != su[:-1]])       # noqa: E731
o = np.lexsort((-rank_i[ti], tu))
key = np.zeros(U, np.int64)
key[tu[o][first(tu[o])]] = rank_i[ti[o][first(tu[o])]]
evaluate(relabel(key)[tu], ti, "users by their coldest item")
for nres in (8000 // scale * 4, 32000 // scale * 4):
    r = rank_i[ti]
    score = np.where(r >= nres, r, r + 10 * I)
    o = np.lexsort((score, tu))
    key = np.zeros(U, np.int64)
    key[tu[o][first(tu[o])]] = rank_i[ti[o][first(tu[o])]]
    evaluate(relabel(key)[tu], ti, f"users by hottest item outside the top {nres}")
