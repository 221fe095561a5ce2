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


def test_batched_transpose_softmax_bwd_lnbwd_x2(dev):
    from mqdet_b200 import ops
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 3, 50, 36, generator=g).half()
    xv = x.permute(0, 2, 1, 3)                                   # strided view [2, 50, 3, 36] -> batch dims (0, 2)
    o = ops.transpose_cast_batched(xv.permute(0, 2, 1, 3).to(dev)).cpu()
    assert o.shape == (2, 3, 36, 56) and torch.equal(o[..., :50], x.transpose(-1, -2)) and (o[..., 50:] == 0).all()
    base = torch.randn(2, 50, 3 * 36, generator=g).half().to(dev)  # [B, T, H*d] read as [B, H, T, d] through strides
    o2 = ops.transpose_cast_batched(base.view(2, 50, 3, 36).permute(0, 2, 1, 3)).cpu()
    assert torch.equal(o2[..., :50], base.cpu().view(2, 50, 3, 36).permute(0, 2, 3, 1))
    s = torch.randn(4, 6, 64, 64, generator=g)
    pr = s.clone().requires_grad_(True)
    p = pr.softmax(-1)
    dp = torch.randn(4, 6, 64, 64, generator=g)
    p.backward(dp)
    ds = ops.softmax_bwd_rows(p.detach().half().to(dev), dp.to(dev), scale=0.5)
    assert_close(ds, 0.5 * pr.grad, 2e-3, "softmax_bwd_rows")
    a, b = torch.randn(300, 768, generator=g), torch.randn(300, 768, generator=g)
    w = torch.randn(768, generator=g)
    dy = torch.randn(300, 768, generator=g)
    ar = (a + b).requires_grad_(True)
    torch.nn.functional.layer_norm(ar, (768,), w, torch.zeros(768), 1e-12).backward(dy)
    dx, _, _ = ops.layernorm_bwd(dy.to(dev), a.to(dev), w.to(dev), 1e-12, want_param_grads=False, x2=b.to(dev))
    assert_close(dx, ar.grad, 1e-5, "layernorm_bwd with two addends")


def test_bert_layer_backward_vs_autograd(dev):
    from mqdet_b200 import ops
    from mqdet_b200.modeling.language_backbone.bert_backward import BertLayerTrain
    from mqdet_b200.modeling.language_backbone.modeling_bert_new import BertLayer
    from oracle import restate, synth
    gen = synth.Gen(8)
    sd = synth.bert_layer_sd(gen, "")
    B, T = 8, 256
    h, dy = gen.randn(B, T, 768), gen.randn(B, T, 768)
    am = torch.ones(B, T)
    am[0, 200:] = 0
    hr = h.clone().requires_grad_(True)
    y = restate.bert_layer(hr, restate.extended_mask(am), sd, "", 12)
    y.backward(dy)
    tr = BertLayerTrain(load_sd(BertLayer(768, 12, 3072), sd).to(dev))
    o32, _ = tr.forward(h.to(dev), ops.cast_f16(h.to(dev)), am.to(dev))
    assert_close(o32, y.detach(), FP16_TOL, "bert train forward")
    assert_close(tr.backward(dy.to(dev)), hr.grad, 3e-3, "bert layer backward: dh")


def test_qvbert_encoder_backward_vs_autograd(dev):
    """Two [GCP block, BERT layer] pairs chained: dL/dh at the entry, dL/d(vision) summed over the blocks and every qv_layer gradient
    against autograd over the oracle — the complete backward of the trainable half of the language backbone, given dL/d(hidden)."""
    from mqdet_b200.modeling.language_backbone.gcp_backward import QVBertEncoderTrain
    from mqdet_b200.modeling.language_backbone.modeling_bert_new import QVBertEncoder
    from oracle import restate, synth
    from types import SimpleNamespace
    gen = synth.Gen(12)
    cfgb = SimpleNamespace(hidden_size=768, num_hidden_layers=3, num_attention_heads=12, intermediate_size=3072, layer_norm_eps=1e-12)
    sd = {}
    for i in range(3):
        synth.bert_layer_sd(gen, f"layer.{i}.", sd=sd)
    for i in range(2):
        synth.gcp_block_sd(gen, f"qv_layer.{i}.", sd=sd)
    B, T = 2, 256
    _, _, pmap = synth.prompt(10, 2, T, gen)
    _, m = synth.vision_queries(pmap, 5, T, 768, gen)
    mask = m.expand(B, -1, -1).clone()
    vision, h, dy = gen.randn(B, mask.shape[1], 768), gen.randn(B, T, 768), gen.randn(B, T, 768)
    am = torch.ones(B, T)
    am[1, 220:] = 0
    p = {k: v.clone().requires_grad_(k.startswith("qv_layer")) for k, v in sd.items()}
    hr, vr = h.clone().requires_grad_(True), vision.clone().requires_grad_(True)
    x = hr
    for i in (1, 2):
        x = restate.gcp_block(x, vr, mask, p, f"qv_layer.{i - 1}.")
        x = restate.bert_layer(x, restate.extended_mask(am), p, f"layer.{i}.", 12)
    x.backward(dy)
    enc = load_sd(QVBertEncoder(cfgb, dim=768, start_qv_layer_index=1, cfg=vq_cfg()), sd).to(dev)
    tr = QVBertEncoderTrain(enc)
    out = tr.forward(h.to(dev), am.to(dev), vision.to(dev), mask.to(dev))
    assert_close(out, x.detach(), 2e-3, "encoder train forward")
    dh, dvis, grads = tr.backward(dy.to(dev))
    errs = []
    assert_close(dh, hr.grad, 5e-3, "encoder backward: dh", defer=errs)
    assert_close(dvis, vr.grad, 5e-3, "encoder backward: dvision", defer=errs)
    for k, g in grads.items():
        # ff_gate is ONE scalar = a sum of B*T*D signed products: fp16 rounding of the operands does not average out against max|ref|
        assert_close(g.view(p[k].shape), p[k].grad, 1e-1 if k.endswith("ff_gate") else 8e-3, f"encoder backward: d {k}", defer=errs)
    assert len(grads) == 32 and not errs, errs


@pytest.mark.parametrize("B,V,I", [(2, 50, 237), (8, 400, 5577)])
def test_preselect_backward_vs_autograd(dev, B, V, I):
    """Gradient of all 23 PreSelect parameter tensors (and of the incoming vision queries) against autograd over the oracle;
    (8, 400, 5577) is the per-GPU shape of BASELINE configs 2 / 5 (80 classes x 5 queries, the pooled 800x1333 pyramid)."""
    from mqdet_b200.modeling.language_backbone.modeling_bert_new import PreSelectModule
    from mqdet_b200.modeling.language_backbone.preselect_backward import PreSelectTrain
    from oracle import restate, synth
    gen = synth.Gen(21)
    sd = synth.preselect_sd(gen)
    vision, image, dy = gen.randn(B, V, 256), gen.randn(B, I, 256), gen.randn(B, V, 768)
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    vr = vision.clone().requires_grad_(True)
    y = restate.preselect(vr, image, p, "")
    y.backward(dy)
    tr = PreSelectTrain(load_sd(PreSelectModule(dim=256, out_dim=768, cfg=vq_cfg()), sd).to(dev))
    out = tr.forward(vision.to(dev), image.to(dev))
    assert_close(out, y.detach(), FP16_TOL, "preselect train forward")
    dvin, grads = tr.backward(dy.to(dev))
    errs = []
    assert_close(dvin, vr.grad, 3e-3, "preselect backward: d vision", defer=errs)
    assert set(grads) == set(sd)
    for k in sd:
        assert_close(grads[k].view(p[k].shape), p[k].grad, 3e-3, f"preselect backward: d {k}", defer=errs)
    assert not errs, errs


def test_qvbert_model_backward_and_optimizer_step(dev):
    """The whole trainable half of the language backbone: QVBertModel forward (12 BERT layers, PreSelect, 6 GCP blocks) -> a token focal
    loss on a linear read-out of the hidden state -> backward to all 119 trainable tensors (encoder.qv_layer.*, pre_select.*) against
    autograd over the oracle, then one FusedAdamW step with the reference's parameter-group rules against torch.optim.AdamW."""
    from types import SimpleNamespace as NS
    from mqdet_b200 import ops
    from mqdet_b200.modeling.language_backbone.bert_model_new import bert_base_config
    from mqdet_b200.modeling.language_backbone.gcp_backward import QVBertModelTrain
    from mqdet_b200.modeling.language_backbone.modeling_bert_new import QVBertModel
    from mqdet_b200.solver.build import FusedAdamW, param_group_options
    from oracle import restate, synth
    gen = synth.Gen(31)
    sd = synth.qvbert_sd(gen)
    B, T, I = 2, 256, 301
    ids, am, pmap = synth.prompt(10, 2, T, gen)
    ids, am = ids.expand(B, -1).contiguous(), am.expand(B, -1).contiguous()
    vision, vmask = synth.vision_queries(pmap, 5, T, 256, gen)
    vision, vmask = vision.expand(B, -1, -1).contiguous(), vmask.expand(B, -1, -1).contiguous()
    images = gen.randn(B, I, 256)
    readout = gen.randn(T, 768, scale=0.05)                       # logits[b, n, t] = hidden[b, n] . readout[t]
    targets = (torch.rand(B, T, T, generator=gen.g) > 0.98).float()
    train_keys = [k for k in sd if k.startswith(("encoder.qv_layer", "pre_select"))]
    p = {k: (v.clone().requires_grad_(True) if k in train_keys else v) for k, v in sd.items()}
    out = restate.qvbert_model(ids, am, vision, images, vmask, p)
    logits = out["hidden"] @ readout.t()
    loss = restate.token_focal_loss(logits, targets, 0.25, 2.0, am.float()) / 16.0
    loss.backward()
    model = load_sd(QVBertModel(bert_base_config(), dim_t=768, dim_v=256, cfg=vq_cfg()), sd).to(dev)
    tr = QVBertModelTrain(model)
    hid = tr.forward(ids.to(dev), am.to(dev), vision.to(dev), images.to(dev), vmask.to(dev))
    assert_close(hid, out["hidden"].detach(), 5e-3, "qvbert train forward")
    r16 = ops.cast_f16(readout.to(dev))
    lg = ops.gemm(ops.cast_f16(hid.contiguous()).view(B * T, 768), r16, out_dtype=torch.float32).view(B, T, T)
    l, dl = ops.token_focal_loss(lg, targets.to(dev), am.float().to(dev), 0.25, 2.0, grad_scale=1.0 / 16.0)
    assert abs(l.item() / 16.0 - loss.item()) <= 2e-2 * abs(loss.item())
    dh = ops.gemm(ops.cast_f16(dl).view(B * T, T), ops.transpose_cast(readout.to(dev)), out_dtype=torch.float32).view(B, T, 768)
    grads = tr.backward(dh)
    assert set(grads) == set(train_keys)
    errs = []
    for k in train_keys:
        tol = 1e-1 if k.endswith("ff_gate") else 2e-2   # up to 6 GCP blocks + 6 BERT layers of fp16-operand products between loss and parameter
        assert_close(grads[k].view(p[k].shape), p[k].grad, tol, f"qvbert backward: d {k}", defer=errs)
    assert not errs, errs[:10]
    # one optimizer step, reference parameter groups
    cfg = NS(SOLVER=NS(BASE_LR=1e-4, WEIGHT_DECAY=1e-4, LANG_LR=1e-5, BACKBONE_BODY_LR_FACTOR=1.0, BIAS_LR_FACTOR=2.0, WEIGHT_DECAY_BIAS=0.0,
                       WEIGHT_DECAY_NORM_FACTOR=1.0, GATE_LR=5e-3, QUERY_LR=1e-5, OPTIMIZER="ADAMW",
                       CLIP_GRADIENTS=NS(ENABLED=True, CLIP_TYPE="full_model", CLIP_VALUE=1.0, NORM_TYPE=2.0)))
    named = [("language_backbone.body.model." + k, q) for k, q in model.named_parameters() if k in train_keys]
    for _, q in named:
        q.requires_grad_(True)
    opt = FusedAdamW(named, cfg=cfg)
    ref_params = [torch.nn.Parameter(p[k].detach().clone()) for k in train_keys]
    groups = []
    for k, q in zip(train_keys, ref_params):
        lr, wd = param_group_options(cfg, "language_backbone.body.model." + k)
        groups.append({"params": [q], "lr": lr, "weight_decay": wd})
        q.grad = p[k].grad.clone()
    ref_opt = torch.optim.AdamW(groups)
    norm = torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
    ref_opt.step()
    coef = opt.step({"language_backbone.body.model." + k: grads[k] for k in train_keys}).cpu()
    assert abs(coef[1].item() - norm.item()) <= 2e-2 * norm.item()
    # AdamW's first update is lr * g / (|g| + eps), i.e. sign-like: elements whose gradient is below the fp16 noise may flip, so the
    # updates are compared as vectors (cosine per tensor) and by their size (= the group's learning rate)
    name_to_param = dict(model.named_parameters())
    worst_cos, worst_size = 1.0, 0.0
    for k, q in zip(train_keys, ref_params):
        u_ref = (q.detach() - sd[k]).flatten()
        u = (name_to_param[k].detach().cpu() - sd[k]).flatten()
        if u.numel() >= 64:
            worst_cos = min(worst_cos, torch.nn.functional.cosine_similarity(u, u_ref, dim=0).item())
        worst_size = max(worst_size, abs(u.abs().max().item() / (u_ref.abs().max().item() + 1e-20) - 1.0))
    print("adamw update agreement: worst cosine", worst_cos, "worst size deviation", worst_size)
    assert worst_cos > 0.8 and worst_size < 0.1, (worst_cos, worst_size)
