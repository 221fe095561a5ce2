"""CPU: the oracle restatement (oracle/restate.py) reproduces the golden vectors that oracle/make_golden.py recorded
from the REFERENCE's own modules (tests/golden/*.pt).  fp32 vs fp32, so the tolerance is summation-order noise."""
import os

import pytest
import torch

from util import ROOT

from oracle import make_golden, restate

GOLD = os.path.join(ROOT, "tests", "golden")


def _cmp(fx, key, full, tol=2e-5):
    steps = fx["subsample"][key]
    got = make_golden.sub(full.float(), *steps)
    ref = fx[key]
    err = (got - ref).abs().max().item()
    bound = tol * fx[key + "_absmax"] + 1e-6
    assert err <= bound, f"{fx['case']}.{key}: {err:.3e} > {bound:.3e}"


def test_gcp_block_golden():
    fx = torch.load(os.path.join(GOLD, "gcp_block.pt"))
    c = make_golden.case_inputs("gcp_block")
    _cmp(fx, "y", restate.gcp_block(c["x"], c["vision"], c["mask"], c["sd"]))
    _cmp(fx, "s", restate.gcp_sparse_attention(c["x"], c["vision"], c["mask"], c["sd"], "attn."))


def test_preselect_golden():
    fx = torch.load(os.path.join(GOLD, "preselect.pt"))
    c = make_golden.case_inputs("preselect")
    _cmp(fx, "vision", restate.preselect(c["vision"], c["image"], c["sd"]))


def test_bi_attention_golden():
    fx = torch.load(os.path.join(GOLD, "bi_attention.pt"))
    c = make_golden.case_inputs("bi_attention")
    v = torch.cat([f.flatten(2).transpose(1, 2) for f in c["feats"]], dim=1)
    v2, l2 = restate.bi_attention(v, c["l"], c["mask"], c["sd"])
    _cmp(fx, "v", v2)
    _cmp(fx, "l", l2)


def test_bert_layer_golden():
    fx = torch.load(os.path.join(GOLD, "bert_layer.pt"))
    c = make_golden.case_inputs("bert_layer")
    h = restate.bert_layer(c["h"], restate.extended_mask(c["am"]), c["sd"], "", clamp=50000.0)
    _cmp(fx, "h", h)
    # the +-5e4 clamps are inactive at these magnitudes: the un-clamped HF-style layer gives the same numbers
    _cmp(fx, "h", restate.bert_layer(c["h"], restate.extended_mask(c["am"]), c["sd"], ""))


def test_gcp_semantic_traps():
    """Properties the reference code implies (SURVEY.md §7 'semantic traps')."""
    c = make_golden.case_inputs("gcp_block")
    sd, x, vision, mask = c["sd"], c["x"], c["vision"], c["mask"]
    s = restate.gcp_sparse_attention(x, vision, mask, sd, "attn.")
    no_cls = mask.sum(1) == 0
    assert s[no_cls].abs().max().item() == 0.0  # all-padding rows: softmax*0 -> exactly zero
    # K/V depend only on the unique query: permuting tokens permutes outputs
    perm = torch.randperm(256, generator=torch.Generator().manual_seed(0))
    s2 = restate.gcp_sparse_attention(x[:, perm], vision, mask[:, :, perm], sd, "attn.")
    assert torch.allclose(s2, s[:, perm], atol=1e-6)
    # identity at the reference's zero-initialised gates (modeling_bert_new.py:274,289)
    sd0 = dict(sd)
    sd0["attn_gate.linear2.weight"] = torch.zeros_like(sd["attn_gate.linear2.weight"])
    sd0["ff_gate"] = torch.zeros(1)
    assert torch.equal(restate.gcp_block(x, vision, mask, sd0), x)


def test_ml_nms_oracle_properties():
    g = torch.Generator().manual_seed(0)
    n = 400
    xy = torch.rand(n, 2, generator=g) * 300
    wh = torch.rand(n, 2, generator=g) * 80 + 5
    boxes = torch.cat([xy, xy + wh], 1)
    scores = torch.rand(n, generator=g)
    labels = torch.randint(1, 4, (n,), generator=g).float()
    keep = restate.ml_nms(boxes, scores, labels, 0.6)
    assert torch.equal(keep, torch.sort(keep)[0])  # ascending original indices (ml_nms.cu:145-149)
    # idempotence: NMS of the kept set keeps everything
    k2 = restate.ml_nms(boxes[keep], scores[keep], labels[keep], 0.6)
    assert k2.numel() == keep.numel()
    # label gating: with all-distinct labels nothing is suppressed
    assert restate.ml_nms(boxes, scores, torch.arange(n).float(), 0.6).numel() == n
    # class-agnostic case equals torchvision-style greedy NMS with the +1 convention when boxes are far apart
    assert restate.ml_nms(boxes[:1], scores[:1], labels[:1], 0.6).tolist() == [0]


def test_swin_fpn_golden():
    """Restated Swin-T + FPN vs vectors recorded from the reference's SwinTransformer / FPN modules."""
    fx = torch.load(os.path.join(GOLD, "swin_fpn.pt"))
    c = make_golden.case_inputs("swin_fpn")
    outs = restate.swin_transformer(c["img"], c["sd"])
    pyr = restate.fpn(outs, c["fsd"])
    for i, o in enumerate(outs[1:]):
        _cmp(fx, f"c{i + 3}", o, 1e-4)
    for i, o in enumerate(pyr):
        _cmp(fx, f"p{i + 3}", o, 1e-4)


def test_dcn_fast_path_equals_gather_formulation():
    """oracle dcn_v2: the torchvision fast path (used when offsets are not re-interpreted) equals the explicit restatement
    of deform_conv_kernel_cuda.cu:578-641, for stride 1 and 2."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 16, 9, 11, generator=g)
    w = torch.randn(8, 16, 3, 3, generator=g) * 0.1
    b = torch.randn(8, generator=g)
    for stride in (1, 2):
        Ho, Wo = (9 - 1) // stride + 1, (11 - 1) // stride + 1
        off = torch.randn(2, 18 * Ho * Wo, generator=g) * 2
        m = torch.rand(2, 9 * Ho * Wo, generator=g)
        a = restate.dcn_v2(x, off, m, w, b, stride, fast=True)
        c = restate.dcn_v2(x, off, m, w, b, stride, fast=False)
        assert (a - c).abs().max().item() <= 1e-5


def test_contrastive_embed_golden():
    """GroundingDINO ContrastiveEmbed: -inf pattern exact, finite logits to summation-order noise."""
    fx = torch.load(os.path.join(GOLD, "contrastive_embed.pt"))
    c = make_golden.case_inputs("contrastive_embed")
    got = make_golden.sub(restate.contrastive_embed(c["x"], c["y"], c["mask"], 256), *fx["subsample"]["logits"])
    ref = fx["logits"]
    assert torch.equal(torch.isinf(got), torch.isinf(ref))
    fin = torch.isfinite(ref)
    assert (got[fin] - ref[fin]).abs().max().item() <= 2e-5 * fx["logits_absmax"]


def test_dyconv_golden():
    fx = torch.load(os.path.join(GOLD, "dyconv.pt"))
    c = make_golden.case_inputs("dyconv")
    _cmp(fx, "v", restate.flatten_levels(restate.dyconv(c["feats"], c["sd"])), 1e-4)


def test_vldyhead_golden():
    """Fixture recorded from the reference's own VLDyHead.forward (DCN kernel substituted): logits, language stream, boxes."""
    fx = torch.load(os.path.join(GOLD, "vldyhead.pt"))
    c = make_golden.case_inputs("vldyhead")
    got = restate.vl_dyhead(c["feats"], c["hidden"], c["masks"], c["sd"])
    _cmp(fx, "logits", got["dot_product_logits"], 2e-4)
    _cmp(fx, "hidden", got["hidden"], 2e-4)
    _cmp(fx, "bbox", restate.flatten_levels(got["bbox_reg"]), 2e-4)
    _cmp(fx, "ctr", restate.flatten_levels(got["centerness"]), 2e-4)


def test_detector_golden():
    """Detections recorded from the reference's own GeneralizedVLRCNN_New.forward (whole model, CPU)."""
    fx = torch.load(os.path.join(GOLD, "detector.pt"))
    c = make_golden.case_inputs("detector")
    got = restate.detector(c["img"], c["size"], c["ids"], c["am"], c["pmap"], c["bank"], c["sd"], K=5, num_classes=80)
    boxes, scores, labels = got["detections"][0]
    have = make_golden.canonical_detections(torch.cat([boxes, scores[:, None], labels[:, None].float()], 1))
    want = fx["det"]
    assert have.shape == want.shape
    assert torch.equal(want[:, 5], have[:, 5])
    assert torch.allclose(want[:, 4], have[:, 4], rtol=2e-4, atol=2e-5) and torch.allclose(want[:, :4], have[:, :4], rtol=0, atol=5e-2)


def test_gdino_transformer_restatement_vs_golden():
    """oracle/restate.py::gdino_transformer against the fixture recorded from the reference's own Transformer.forward."""
    c = make_golden.case_inputs("gdino_transformer")
    poss = [restate.position_embedding_sine_hw(m) for m in c["masks"]]
    out = restate.gdino_transformer(c["srcs"], c["masks"], poss, c["enc_text"], c["tmask"], c["pid"], c["sam"], c["sd"],
                                    num_queries=c["nq"], enc_layers=c["enc_layers"], dec_layers=c["dec_layers"])
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "gdino_transformer.pt"))
    for got, key in ((out["hs"][-1], "hs_last"), (out["references"][-1], "ref_last"), (out["memory_text"], "text")):
        assert (got - fx[key]).abs().max().item() <= 2e-5 * fx[key + "_absmax"] + 1e-6
