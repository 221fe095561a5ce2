import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mqdet_b200 import ops
dev = torch.device("cuda:0")
for (M, N, ld, dt) in [(8, 21, 24, torch.float16), (8, 21, 24, torch.float32), (130, 70, 72, torch.float16), (8, 5577, 5584, torch.float16)]:
    a = torch.ones(M, 64, device=dev).half()
    b = torch.ones(N, 64, device=dev).half()
    out = torch.full((M, ld), -7.0, dtype=dt, device=dev)
    ops.gemm(a, b, out=out[:, :N])
    torch.cuda.synchronize()
    bad = (out[:, N:] != -7.0)
    print(M, N, ld, dt, 'pad touched:', int(bad.sum()), 'cols', sorted(set((bad.nonzero()[:, 1] + N).tolist()))[:10], 'in-bounds ok:', bool((out[:, :N] == 64).all()))
    print(' row0 tail', out[0, max(0, N - 3):].tolist())
