"""The encoder-shaped multi-scale deformable attention core alone (2 images, 22323 queries, 8 heads, 4 levels x 4 points; BASELINE
config 4), for ncu captures and a CUDA-event timing with the algorithmic gather rate:

    python tools/prof_msda.py [reps]
    ncu --set full --import-source on --clock-control none -k regex:ms_deform_attn -c 1 --launch-skip 2 -o gpurun_out/x python tools/prof_msda.py 3
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from mqdet_b200 import ops

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 7
g = torch.Generator().manual_seed(0)
sizes = [(100, 168), (50, 84), (25, 42), (13, 21)]
B, H, L, P = 2, 8, 4, 4
N = sum(h * w for h, w in sizes)
levels = ops.get_levels(sizes, dev)
value = (torch.randn(B, N, 256, generator=g)).half().to(dev)
proj = torch.cat([torch.randn(B * N, H * L * P * 2, generator=g) * 2.0, torch.randn(B * N, H * L * P, generator=g)], 1).to(dev).contiguous()
ref = torch.rand(B, N, L, 2, generator=g).to(dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ts = []
for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops.ms_deform_attn(value, proj, H * L * P * 2, ref, levels, H, P)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
ms = ts[len(ts) // 2]
gather = B * N * H * L * P * 4 * 64          # corner reads: 64 bytes per (query, head, sample, corner)
alg = value.numel() * 2 + proj.numel() * 4 + ref.numel() * 4 + out.numel() * 2   # every operand once
print(f"ms_deform_attn encoder shape: {ms:.3f} ms; gathers {gather / 1e9:.2f} GB -> {gather / ms / 1e6:.0f} GB/s from L2; "
      f"algorithmic HBM bytes {alg / 1e6:.1f} MB -> {alg / ms / 1e6:.0f} GB/s")
