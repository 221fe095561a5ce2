"""GPU parity of the training-side kernels (BASELINE config 5) through the C ABI: the GCP block's backward against
torch.autograd over the CPU oracle (oracle/restate.py::gcp_block, pinned to the reference's GatedCrossAttentionBlock), the token focal
loss against the restated reference formula (pinned in tests/test_train_cpu.py), AdamW + global-norm clipping against torch.optim."""
import math

import pytest
import torch

from util import FP16_TOL, assert_close, load_sd, vq_cfg

pytestmark = pytest.mark.gpu


def test_transpose_cast(dev):
    from mqdet_b200 import ops
    g = torch.Generator().manual_seed(1)
    for dt in (torch.float32, torch.float16):
        x = torch.randn(203, 77, generator=g).to(dt)
        o = ops.transpose_cast(x.to(dev), scale=0.5).cpu()
        assert o.shape == (77, 208)
        assert torch.equal(o[:, :203], (x.float() * 0.5).t().half()) and (o[:, 203:] == 0).all()


def test_layernorm_bwd_vs_autograd(dev):
    from mqdet_b200 import ops
    g = torch.Generator().manual_seed(2)
    for rows, D in ((2048, 768), (403, 768), (37, 256)):
        x = torch.randn(rows, D, generator=g) * 2 + 0.5
        w, b = torch.randn(D, generator=g), torch.randn(D, generator=g)
        dy = torch.randn(rows, D, generator=g)
        xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        torch.nn.functional.layer_norm(xr, (D,), wr, br, 1e-5).backward(dy)
        base = torch.randn(rows, D, generator=g)
        dx, dg, db = ops.layernorm_bwd(dy.to(dev), x.to(dev), w.to(dev), 1e-5, dx=base.to(dev).clone())
        assert_close(dx, base + xr.grad, 1e-5, f"layernorm_bwd dx (+=) {rows}x{D}")
        assert_close(dg, wr.grad, 1e-5, f"layernorm_bwd dgamma {rows}x{D}")
        assert_close(db, br.grad, 1e-5, f"layernorm_bwd dbeta {rows}x{D}")
        dx2, _, _ = ops.layernorm_bwd(dy.to(dev), x.to(dev), w.to(dev), 1e-5, want_param_grads=False)
        assert_close(dx2, xr.grad, 1e-5, f"layernorm_bwd dx (=) {rows}x{D}")


def test_gelu_bwd_vs_autograd(dev):
    from mqdet_b200 import ops
    g = torch.Generator().manual_seed(3)
    z = (torch.randn(1000, 384, generator=g) * 2).half()
    for dh in (torch.randn(1000, 384, generator=g), torch.randn(1000, 384, generator=g).half()):
        zr = z.float().requires_grad_(True)
        torch.nn.functional.gelu(zr).backward(dh.float())
        assert_close(ops.gelu_bwd(z.to(dev), dh.to(dev)), zr.grad, FP16_TOL, f"gelu_bwd {dh.dtype}")


def test_token_focal_loss_vs_oracle(dev):
    from mqdet_b200 import ops
    from oracle import restate
    g = torch.Generator().manual_seed(4)
    B, N, T = 2, 1500, 256
    logits = torch.randn(B, N, T, generator=g) * 3 - 2
    targets = (torch.rand(B, N, T, generator=g) > 0.97).float()
    tm = torch.ones(B, T)
    tm[0, 200:] = 0
    tm[1, 120:] = 0
    for mask in (tm, None):
        lr = logits.clone().requires_grad_(True)
        ref = restate.token_focal_loss(lr, targets, 0.25, 2.0, mask)
        ref.backward()
        loss, dl = ops.token_focal_loss(logits.to(dev), targets.to(dev), None if mask is None else mask.to(dev), 0.25, 2.0)
        assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
        assert_close(dl, lr.grad, 1e-5, "token_focal_loss dlogits")
    _, dl2 = ops.token_focal_loss(logits.to(dev), targets.to(dev), tm.to(dev), 0.25, 2.0, grad_scale=1024.0)
    lr = logits.clone().requires_grad_(True)
    restate.token_focal_loss(lr, targets, 0.25, 2.0, tm).backward()
    assert_close(dl2, lr.grad * 1024.0, 1e-5, "token_focal_loss dlogits with loss scaling")


def test_adamw_and_clipping_vs_torch(dev):
    from mqdet_b200.solver.build import FusedAdamW
    g = torch.Generator().manual_seed(5)
    shapes = [(768, 3072), (768,), (1,), (384, 768)]
    ps = [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes]
    ref_ps = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt_ref = torch.optim.AdamW([{"params": [p], "lr": 1e-3 * (i + 1), "weight_decay": 0.01 * i} for i, p in enumerate(ref_ps)])
    mine = [torch.nn.Parameter(p.detach().clone().to(dev)) for p in ps]
    opt = FusedAdamW([(f"p{i}", p) for i, p in enumerate(mine)], clip_value=1.0)
    for i, st in enumerate(opt.params.values()):
        st["lr"], st["wd"] = 1e-3 * (i + 1), 0.01 * i
    for step in range(3):
        grads = [torch.randn(*s, generator=g) * (5.0 if step == 0 else 0.01) for s in shapes]
        for p, gr in zip(ref_ps, grads):
            p.grad = gr.clone()
        norm = torch.nn.utils.clip_grad_norm_(ref_ps, 1.0)
        opt_ref.step()
        coef = opt.step({f"p{i}": gr.to(dev) for i, gr in enumerate(grads)}).cpu()
        assert abs(coef[1].item() - norm.item()) <= 1e-4 * norm.item()
        assert abs(coef[0].item() - min(1.0, 1.0 / (norm.item() + 1e-6))) <= 1e-5
        for a, b in zip(mine, ref_ps):
            assert_close(a.detach(), b.detach(), 1e-5, f"adamw parameter after step {step + 1}")


def _gcp_case(seed, B, T, ncls):
    from oracle import synth
    gen = synth.Gen(seed)
    sd = synth.gcp_block_sd(gen)
    _, _, pmap = synth.prompt(ncls, 2, T, gen)
    _, m = synth.vision_queries(pmap, 5, T, 768, gen)
    mask = m.expand(B, -1, -1).clone()
    mask[0, 3] = 0
    mask[B - 1, 7:10] = 0
    V = mask.shape[1]
    return sd, mask, gen.randn(B, V, 768), gen.randn(B, T, 768), gen.randn(B, T, 768)


@pytest.mark.parametrize("B,T,ncls", [(2, 256, 10), (8, 256, 80)])
def test_gcp_block_backward_vs_autograd(dev, B, T, ncls):
    """dx, dvision and the gradient of every parameter of a GatedCrossAttentionBlock against torch.autograd over the fp32 CPU oracle;
    (8, 256, 80) is the per-GPU shape of the pre-training step (BASELINE config 5: 8 images / GPU, <= 85 classes)."""
    from mqdet_b200.modeling.language_backbone.gcp_backward import GCPBlockTrain
    from mqdet_b200.modeling.language_backbone.modeling_bert_new import GatedCrossAttentionBlock
    from oracle import restate
    sd, mask, vision, x, dy = _gcp_case(90 + B, B, T, ncls)
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr, vr = x.clone().requires_grad_(True), vision.clone().requires_grad_(True)
    y = restate.gcp_block(xr, vr, mask, p, "")
    y.backward(dy)
    blk = load_sd(GatedCrossAttentionBlock(dim=768, cfg=vq_cfg()), sd).to(dev)
    tr = GCPBlockTrain(blk)
    yy = tr.forward(x.to(dev), vision.to(dev), mask.to(dev))
    assert_close(yy, y.detach(), FP16_TOL, "gcp train forward")
    dx, dvis, grads = tr.backward(dy.to(dev))
    errs = []
    tol = 3e-3   # fp16 operands through up to five chained products (dy -> dz2 -> dx1 -> ds -> do -> dq/dkv -> dW)
    assert_close(dx, xr.grad, tol, "gcp backward: dx", defer=errs)
    assert_close(dvis, vr.grad, tol, "gcp backward: dvision", defer=errs)
    assert set(grads) == set(sd)
    for k in sd:
        assert_close(grads[k].view(p[k].shape), p[k].grad, tol, f"gcp backward: d {k}", defer=errs)
    assert not errs, errs


def test_gcp_training_step_reduces_loss(dev):
    """Three optimizer steps on one GCP block (forward -> loss gradient -> backward -> clip -> AdamW, all on the device) move the output
    towards a target: the pieces compose into a working update."""
    from mqdet_b200.modeling.language_backbone.gcp_backward import GCPBlockTrain
    from mqdet_b200.modeling.language_backbone.modeling_bert_new import GatedCrossAttentionBlock
    from mqdet_b200.solver.build import FusedAdamW
    sd, mask, vision, x, _ = _gcp_case(77, 2, 256, 10)
    blk = load_sd(GatedCrossAttentionBlock(dim=768, cfg=vq_cfg()), sd).to(dev)
    for p_ in blk.parameters():
        p_.requires_grad_(True)
    target = torch.zeros(2, 256, 768, device=dev)
    opt = FusedAdamW(list(blk.named_parameters()), lr=1e-3, weight_decay=0.0, clip_value=1.0)
    tr = GCPBlockTrain(blk)
    losses = []
    xd, vd, md = x.to(dev), vision.to(dev), mask.to(dev)
    for _ in range(4):
        y = tr.forward(xd, vd, md)
        losses.append(((y - xd - target) ** 2).mean().item())   # pull the block's contribution (y - x) to zero
        dy = 2.0 * (y - xd - target) / y.numel()
        _, _, grads = tr.backward(dy * 1000.0)
        opt.step(grads)
    assert all(b < a for a, b in zip(losses, losses[1:])) and losses[-1] < losses[0] * 0.97, losses
