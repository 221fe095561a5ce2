"""CPU, build container only: pins oracle/restate.py against the reference's own modules executed from
/root/reference (full tensors, not the committed subsamples).  Skipped where the reference is absent (GPU box)."""
import pytest
import torch

from oracle import make_golden, ref_loader, restate

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


def _close(a, b, tol=2e-5):
    err = (a - b).abs().max().item()
    assert err <= tol * b.abs().max().item() + 1e-6, f"{err:.3e}"


def test_gcp_block_vs_reference():
    c = make_golden.case_inputs("gcp_block")
    ref = make_golden.run_reference("gcp_block")
    _close(restate.gcp_block(c["x"], c["vision"], c["mask"], c["sd"]), ref["y"])
    _close(restate.gcp_sparse_attention(c["x"], c["vision"], c["mask"], c["sd"], "attn."), ref["s"])
    # the index table equals the reference's topk-based construction (modeling_bert_new.py:40-63)
    m = ref_loader.modeling_bert_new()
    idx_ref = m.get_index_with_padding_batch(c["mask"].transpose(2, 1))
    assert torch.equal(restate.gcp_index(c["mask"]), idx_ref)


def test_preselect_vs_reference():
    c = make_golden.case_inputs("preselect")
    _close(restate.preselect(c["vision"], c["image"], c["sd"]), make_golden.run_reference("preselect")["vision"])


def test_bi_attention_vs_reference():
    c = make_golden.case_inputs("bi_attention")
    ref = make_golden.run_reference("bi_attention")
    v = torch.cat([f.flatten(2).transpose(1, 2) for f in c["feats"]], dim=1)
    v2, l2 = restate.bi_attention(v, c["l"], c["mask"], c["sd"])
    _close(v2, ref["v"])
    _close(l2, ref["l"])


def test_bert_layer_vs_reference():
    c = make_golden.case_inputs("bert_layer")
    ref = make_golden.run_reference("bert_layer")
    _close(restate.bert_layer(c["h"], restate.extended_mask(c["am"]), c["sd"], "", clamp=50000.0), ref["h"])


def test_swin_fpn_vs_reference():
    c = make_golden.case_inputs("swin_fpn")
    ref = make_golden.run_reference("swin_fpn")
    outs = restate.swin_transformer(c["img"], c["sd"])
    for i, o in enumerate(outs):
        _close(o, ref[f"c{i + 2}"], 1e-5)
    for i, o in enumerate(restate.fpn(outs, c["fsd"])):
        _close(o, ref[f"p{i + 3}"], 1e-5)


def test_contrastive_embed_vs_reference():
    c = make_golden.case_inputs("contrastive_embed")
    ref = make_golden.run_reference("contrastive_embed")["logits"]
    got = restate.contrastive_embed(c["x"], c["y"], c["mask"], 256)
    assert got.shape == ref.shape == (2, 900, 256)
    assert torch.equal(torch.isinf(got), torch.isinf(ref))          # padding tokens and columns T..255 are -inf
    fin = torch.isfinite(ref)
    assert torch.equal(got[fin], ref[fin])                           # same fp32 product
