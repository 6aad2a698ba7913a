"""Profiling driver: fusion projection forward/backward, tensor-core vs SIMT."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_recommendation_b200 import _lib, ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_400_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
c = 768
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
E = torch.randn((n, d), device=dev, generator=g) * 0.05
C = torch.randn((n, c), device=dev, generator=g)
W = torch.randn((d, d + c), device=dev, generator=g) * 0.03
b = torch.randn((d,), device=dev, generator=g) * 0.1
gH = torch.randn((n, d), device=dev, generator=g)
H = torch.empty((n, d), device=dev)
lib = _lib.load()


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


flops = 2.0 * n * (d + c) * d
for simt in (0, 1):
    lib.lgcn_fusion_force_simt(simt)
    ms = timeit(lambda: ops.fusion_proj_fwd(E, C, W, b, out=H))
    by = 4.0 * n * (c + 2 * d)
    print(f"fwd {'simt' if simt else 'tc  '} n={n} d={d}: {ms:.2f} ms  {flops / ms / 1e9:.1f} TFLOP/s (fp32-equivalent)  "
          f"{by / ms / 1e6:.0f} GB/s algorithmic")
    href = H.clone() if simt else None
    if not simt:
        Htc = H.clone()
err = (Htc - href).abs().max().item() / href.abs().max().item()
print(f"tc vs simt max rel err {err:.2e}")
gE = torch.empty((n, d), device=dev)
gW = torch.zeros((d, d + c), device=dev)
gb = torch.zeros((d,), device=dev)
for simt in (0, 1):
    lib.lgcn_fusion_force_simt(simt)
    ms = timeit(lambda: ops.fusion_proj_bwd(E, C, W, H, gH, g_eid=gE, gW=gW, gb=gb))
    print(f"bwd {'simt' if simt else 'tc  '} (gE_id + gW + gb) n={n} d={d}: {ms:.2f} ms  "
          f"{(flops + 2.0 * n * d * d) / ms / 1e9:.1f} TFLOP/s (fp32-equivalent)")
    if simt:
        print(f"gE_id tc vs simt max rel err {(gEtc - gE).abs().max().item() / gE.abs().max().item():.2e}")
    else:
        gEtc = gE.clone()
lib.lgcn_fusion_force_simt(0)
