"""GPU parity of the Swin-T backbone (window attention with padding / cyclic shift / region mask in-kernel) and the FPN
against the CPU oracle and the vectors recorded from the reference's SwinTransformer / FPN modules."""
import os

import pytest
import torch

from util import FP16_TOL, ROOT, assert_close, load_sd

pytestmark = pytest.mark.gpu


def _load_swin(sd, dev):
    from mqdet_b200.modeling.backbone.swint import SwinTransformer
    m = SwinTransformer()
    own = m.state_dict()
    full = dict(sd)
    for k in own:
        if k.endswith("relative_position_index"):
            full[k] = own[k]
    return load_sd(m, full).to(dev).eval()


def test_swin_block_shift_and_padding(dev):
    """One shifted block on a 38x51 grid (pads to 42x56): exercises padded keys (= qkv bias), roll and the -100 mask."""
    from mqdet_b200.modeling.backbone.swint import SwinTransformerBlock
    from oracle import restate, synth
    gen = synth.Gen(91)
    sd = {k[len("layers.0.blocks.1."):]: v for k, v in synth.swin_sd(gen).items() if k.startswith("layers.0.blocks.1.")}
    B, H, W, C = 2, 38, 51, 96
    x = gen.randn(B, H * W, C)
    for shift in (0, 3):
        ref = restate.swin_block(x, H, W, sd, "", 3, 7, shift)
        blk = SwinTransformerBlock(C, 3, 7, shift)
        full = dict(sd)
        full["attn.relative_position_index"] = blk.state_dict()["attn.relative_position_index"]
        blk = load_sd(blk, full).to(dev).eval()
        blk.H, blk.W = H, W
        out = blk(x.to(dev), None)
        assert_close(out, ref, what=f"Swin block shift={shift}")


@pytest.mark.parametrize("H,W", [(40, 53), (24, 24), (7, 30)])
def test_swin_l_block_window12(dev, H, W):
    """Swin-L geometry (MQ-GLIP-L, configs/pretrain/mq-glip-l.yaml:11-17): window 12 (144-token windows), 6 heads x 32 at
    embed 192; plain and shifted (6) blocks on grids that need padding, incl. one smaller than the window."""
    from mqdet_b200.modeling.backbone.swint import SwinTransformerBlock
    from oracle import restate, synth
    gen = synth.Gen(92)
    C, heads, ws = 192, 6, 12
    full_sd = synth.swin_sd(gen, depths=(2,), heads=(heads,), embed=C, ws=ws)
    sd = {k[len("layers.0.blocks.1."):]: v for k, v in full_sd.items() if k.startswith("layers.0.blocks.1.")}
    B = 2
    x = gen.randn(B, H * W, C)
    for shift in (0, 6):
        ref = restate.swin_block(x, H, W, sd, "", heads, ws, shift)
        blk = SwinTransformerBlock(C, heads, ws, shift)
        full = dict(sd)
        full["attn.relative_position_index"] = blk.state_dict()["attn.relative_position_index"]
        blk = load_sd(blk, full).to(dev).eval()
        blk.H, blk.W = H, W
        out = blk(x.to(dev), None)
        assert_close(out, ref, what=f"Swin-L block {H}x{W} shift={shift}")


def test_swin_l_backbone(dev):
    """Whole Swin-L body (depths 2,2,18,2; heads 6,12,24,48; window 12) against the oracle on a small image."""
    from mqdet_b200.modeling.backbone.swint import SwinTransformer
    from oracle import restate, synth
    gen = synth.Gen(93)
    depths, heads, embed, ws = (2, 2, 18, 2), (6, 12, 24, 48), 192, 12
    sd = synth.swin_sd(gen, depths=depths, heads=heads, embed=embed, ws=ws)
    img = gen.randn(1, 3, 150, 203)
    ref = restate.swin_transformer(img, sd, depths=depths, heads=heads, embed=embed, ws=ws)
    m = SwinTransformer(embed_dim=embed, depths=depths, num_heads=heads, window_size=ws)
    full = dict(sd)
    for k, v in m.state_dict().items():
        if k.endswith("relative_position_index"):
            full[k] = v
    m = load_sd(m, full).to(dev).eval()
    outs = m(img.to(dev))
    bad = []
    for i, (o, r) in enumerate(zip(outs, ref)):
        assert_close(o, r, 5e-3, f"Swin-L stage{i + 2}", defer=bad)  # 24 blocks with fp16 GEMM operands, fp32 residual
    assert not bad, bad


def test_swin_fpn_vs_oracle_and_golden(dev):
    from mqdet_b200.modeling.backbone.fpn import FPN, LastLevelP6P7
    from oracle import make_golden, restate
    c = make_golden.case_inputs("swin_fpn")
    body = _load_swin(c["sd"], dev)
    fpn = load_sd(FPN([0, 192, 384, 768], 256, LastLevelP6P7(256, 256)), c["fsd"]).to(dev).eval()
    ref_c = restate.swin_transformer(c["img"], c["sd"])
    ref_p = restate.fpn(ref_c, c["fsd"])
    outs = body(c["img"].to(dev))
    assert len(outs) == 4
    bad = []
    for i, (o, r) in enumerate(zip(outs, ref_c)):
        assert_close(o, r, 3e-3, f"Swin stage{i + 2}", defer=bad)  # 12 blocks with fp16 GEMM operands, fp32 residual
    pyr = fpn(outs)
    for i, (o, r) in enumerate(zip(pyr, ref_p)):
        assert_close(o, r, 4e-3, f"FPN P{i + 3}", defer=bad)
    # flat fast path == NCHW API
    feats = body.forward_flat(c["img"].to(dev))
    p16, levels = fpn.forward_flat([feats[i] for i in (1, 2, 3)])
    assert levels.sizes == [tuple(r.shape[2:]) for r in ref_p]
    assert_close(p16, restate.flatten_levels(ref_p), 4e-3, "FPN flat pyramid", defer=bad)
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "swin_fpn.pt"))
    for key, got in [("c3", outs[1]), ("c5", outs[3]), ("p3", pyr[0]), ("p7", pyr[4])]:
        g = make_golden.sub(got.float().cpu(), *fx["subsample"][key])
        err = (g - fx[key]).abs().max().item()
        if err > 4e-3 * fx[key + "_absmax"] + 4e-3:
            bad.append(f"golden {key}: {err:.3e}")
    assert not bad, bad


@pytest.mark.parametrize("C,H,W", [(96, 13, 10), (192, 7, 9), (384, 5, 4), (64, 6, 7)])
def test_patch_merge_ln(dev, C, H, W):
    """PatchMerging gather + LayerNorm(4C) (swin_transformer.py:256-284) incl. the zero padding of odd H / W; C = 96 / 192 /
    384 take the vectorised register-cached kernel, other widths the generic one."""
    from mqdet_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    B = 2
    x = torch.randn(B, H * W, C, generator=g) * 1.5 + 0.3
    gamma, beta = torch.randn(4 * C, generator=g), torch.randn(4 * C, generator=g)
    xi = x.view(B, H, W, C)
    xi = torch.nn.functional.pad(xi, (0, 0, 0, W % 2, 0, H % 2))
    cat = torch.cat([xi[:, 0::2, 0::2], xi[:, 1::2, 0::2], xi[:, 0::2, 1::2], xi[:, 1::2, 1::2]], -1)
    ref = torch.nn.functional.layer_norm(cat, (4 * C,), gamma, beta, 1e-5).reshape(-1, 4 * C)
    out, H2, W2 = ops.patch_merge_ln(x.to(dev).contiguous(), B, H, W, gamma.to(dev), beta.to(dev), 1e-5)
    assert (H2, W2) == ((H + 1) // 2, (W + 1) // 2)
    assert_close(out, ref, 1e-3, "patch_merge_ln")


@pytest.mark.parametrize("H,W", [(32, 448), (30, 301), (7, 5), (64, 1344)])
def test_patchify4(dev, H, W):
    """PatchEmbed input gather (swint.py:413-418): [B,3,H,W] fp32 -> fp16 [B*ceil(H/4)*ceil(W/4), 48], k = c*16 + i*4 + j, zero
    padding on the right / bottom; widths that are / are not multiples of 4 and of the 64-patch run of a CTA."""
    from mqdet_b200 import ops
    g = torch.Generator().manual_seed(H * 1000 + W)
    B = 2
    img = torch.randn(B, 3, H, W, generator=g)
    out, Hp, Wp = ops.patchify4(img.to(dev))
    assert (Hp, Wp) == ((H + 3) // 4, (W + 3) // 4)
    pad = torch.nn.functional.pad(img, (0, Wp * 4 - W, 0, Hp * 4 - H))
    ref = pad.view(B, 3, Hp, 4, Wp, 4).permute(0, 2, 4, 1, 3, 5).reshape(B * Hp * Wp, 48).half()
    assert torch.equal(out.cpu(), ref)


def test_avgpool_levels(dev):
    from mqdet_b200 import ops
    g = torch.Generator().manual_seed(4)
    sizes = [(19, 26), (10, 13), (5, 7), (3, 4), (2, 2)]
    feats = [torch.randn(2, 256, h, w, generator=g) for h, w in sizes]
    lv = ops.Levels(sizes, dev)
    x16 = torch.cat([f.flatten(2).transpose(1, 2) for f in feats], 1).half().to(dev).contiguous()
    out = ops.avgpool2_levels(x16, lv)
    ref = torch.cat([torch.nn.functional.avg_pool2d(f.half().float(), 2).flatten(2).transpose(1, 2) for f in feats], 1)
    assert out.shape == ref.shape
    assert_close(out, ref, 1e-5, "AvgPool2d(2) + flatten + cat")
