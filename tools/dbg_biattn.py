"""Bring-up diagnostics of the fused BiAttention kernels on the GPU box: each kernel against plain torch fp32 math on the
same fp16 operands (not a test; prints per-stage errors)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mqdet_b200 import ops  # noqa: E402


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    for (B, N, T, H) in [(1, 128, 256, 1), (1, 256, 256, 2), (2, 1000, 256, 8), (1, 333, 64, 8)]:
        E = H * 256
        q = (torch.randn(B, N, E, generator=g) * 0.25).half().to(dev)
        k = (torch.randn(B, T, E, generator=g) * 0.25).half().to(dev)
        vlT = (torch.randn(B, E, T, generator=g)).half().to(dev)
        w = (torch.randn(256, E, generator=g) * 0.05).half().to(dev)
        bias = torch.randn(256, generator=g).to(dev)
        gamma = (torch.rand(256, generator=g) + 0.5).to(dev)
        res = torch.randn(B, N, 256, generator=g).half().to(dev)
        vn = torch.randn(B, N, 256, generator=g).half().to(dev)
        mask = torch.ones(B, T)
        mask[0, T // 2:] = 0
        mask = mask.to(dev)
        vl_h = vlT.float().view(B, H, 256, T)                                 # [B,H,d,T]
        mT = torch.einsum("ohd,bhdt->bhot", w.float().view(256, H, 256), vl_h).half().contiguous()
        out, colmax = ops.biattn_image(q, k, mT, bias, gamma, res, mask, 50000.0, H)
        torch.cuda.synchronize()
        qf, kf = q.float().view(B, N, H, 256).permute(0, 2, 1, 3), k.float().view(B, T, H, 256).permute(0, 2, 1, 3)
        S = (qf @ kf.transpose(-1, -2)).clamp(-5e4, 5e4)                     # [B,H,N,T]
        cm_ref = S.max(dim=2)[0].reshape(B * H, T)
        P = torch.softmax(S + torch.where(mask[:, None, None, :] == 0, -9e15, 1.0), dim=-1)
        vl = vlT.float().view(B, H, 256, T).transpose(-1, -2)                 # [B,H,T,256]
        # probabilities travel as fp16; the value and output projections are folded into mT (fp16)
        D = torch.einsum("bhnt,bhot->bno", P.half().float(), mT.float())
        ref = res.float() + gamma * (D + bias)
        print(f"B={B} N={N} T={T} H={H}: image out rel {rel(out, ref):.3e}  colmax abs {float((colmax - cm_ref).abs().max()):.3e}")
        u = torch.empty((B, H, T, 256), dtype=torch.float16, device=dev)
        ops.biattn_text_vn(k.view(B, T, H, 256).permute(0, 2, 1, 3), q.view(B, N, H, 256).permute(0, 2, 1, 3), vn, colmax, 50000.0, u)
        torch.cuda.synchronize()
        Pl = torch.softmax(S.transpose(-1, -2), dim=-1)                       # [B,H,T,N]
        u_ref = Pl @ vn.float()[:, None]
        print(f"    text U rel {rel(u, u_ref):.3e}")


if __name__ == "__main__":
    main()
