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
        vn = torch.randn(B, N, 256, generator=g).half().to(dev)
        gT = (torch.randn(B, H, T, 256, generator=g) * 0.2).half().to(dev)        # folded query/key operand
        gbias = torch.zeros(B, H, T, 8)
        gbias[..., 0] = torch.randn(B, H, T, generator=g)
        gbias = gbias.to(dev)
        mT = (torch.randn(B, H, 256, T, generator=g) * 0.2).half().to(dev)         # folded value/output operand
        bias = torch.randn(256, generator=g).to(dev)
        gamma = (torch.rand(256, generator=g) + 0.5).to(dev)
        res = torch.randn(B, N, 256, generator=g).half().to(dev)
        mask = torch.ones(B, T)
        mask[0, T // 2:] = 0
        mask = mask.to(dev)
        out, colmax = ops.biattn_image(vn, gT, gbias, mT, bias, gamma, res, mask, 50000.0, H)
        torch.cuda.synchronize()
        S = (torch.einsum("bnc,bhtc->bhnt", vn.float(), gT.float()) + gbias[..., 0][:, :, None, :]).clamp(-5e4, 5e4)
        cm_ref = S.max(dim=2)[0].reshape(B * H, T)
        P = torch.softmax(S + torch.where(mask[:, None, None, :] == 0, -9e15, 1.0), dim=-1)
        D = torch.einsum("bhnt,bhot->bno", P.half().float(), mT.float())           # probabilities travel as fp16
        ref = res.float() + gamma * (D + bias)
        print(f"B={B} N={N} T={T} H={H}: image out rel {rel(out, ref):.3e}  colmax abs {float((colmax - cm_ref).abs().max()):.3e}")
        u = torch.empty((B, H, T, 256), dtype=torch.float16, device=dev)
        ops.biattn_text_vn(gT, vn.view(B, 1, N, 256).expand(B, H, N, 256), vn, colmax, 50000.0, u, rowbias=gbias)
        torch.cuda.synchronize()
        Pl = torch.softmax(S.transpose(-1, -2), dim=-1)                            # [B,H,T,N]
        u_ref = Pl @ vn.float()[:, None]
        print(f"    text U rel {rel(u, u_ref):.3e}")


if __name__ == "__main__":
    main()
