"""Host-logic dry runs (CPU, throw-away process): the drop-in modules composed over tests/hostlogic/ops_double.py, compared with the oracle.

    python tests/hostlogic/run.py transformer|model|extract_query|ref_signatures|biattn_split|gcp_bwd|bert_bwd|preselect_bwd|lang_train

Exit code 0 and a line ``PASS <case>`` on success.  The numbers only say that the HOST logic (views, strides, masks, caches, the order and
the operands of every product) is right; kernel parity is the job of the ``-m gpu`` tests.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import ops_double  # noqa: F401,E402  (patches mqdet_b200.ops; must come first)
import torch  # noqa: E402

import util  # noqa: E402
from mqdet_b200 import ops  # noqa: E402
from oracle import restate  # noqa: E402

CPU = torch.device("cpu")


def use_oracle_backbones():
    """Swin / GCP / PreSelect are GPU-validated modules whose kernels (window attention, sparse attention) have no stand-in: for the
    whole-model run they are replaced by the oracle so that the NEW glue around them is what gets exercised."""
    from mqdet_b200.modeling.backbone.swint import SwinTransformer
    from mqdet_b200.modeling.language_backbone import modeling_bert_new as mbn

    def swin_forward_flat(self, img, want=(1, 2, 3)):
        outs = restate.swin_transformer(img, dict(self.state_dict()))
        return {i: (outs[i].flatten(2).transpose(1, 2).contiguous().half(), outs[i].shape[2], outs[i].shape[3]) for i in want}

    from mqdet_b200.modeling.query_selector.query_selector import QuerySelector
    _set = QuerySelector.set_query_bank

    def set_bank_on_cpu(self, bank):
        self.device = "cpu"
        return _set(self, bank)

    QuerySelector.set_query_bank = set_bank_on_cpu
    SwinTransformer.forward_flat = swin_forward_flat
    mbn.GatedCrossAttentionBlock.forward = lambda self, x, vision, attention_mask=None, bp=None: restate.gcp_block(
        x.float(), vision, attention_mask, dict(self.state_dict()), "")
    mbn.PreSelectModule.forward = lambda self, vision, image: {"vision": restate.preselect(vision, image, dict(self.state_dict()), ""),
                                                              "image": image}


def case_transformer():
    import test_gdino_asm_gpu as t
    t.test_transformer_small_vs_oracle_and_golden(CPU)


def case_model():
    import test_gdino_asm_gpu as t
    use_oracle_backbones()
    t.test_groundingdino_forward_vs_oracle(CPU)


def case_extract_query():
    import test_gdino_asm_gpu as t
    use_oracle_backbones()
    t.test_groundingdino_extract_query_vs_oracle(CPU)


def case_biattn_split():
    from mqdet_b200.modeling.groundingdino.fuse_modules import BiAttentionBlock
    from oracle import synth
    gen = synth.Gen(3)
    sd = synth.bi_attention_sd(gen, "", 256, 256, 1024, 6, {})
    B, N, T = 2, 8203, 64
    v, l = gen.randn(B, N, 256), gen.randn(B, T, 256)
    mv = torch.zeros(B, N, dtype=torch.bool)
    mv[1, 8000:] = True
    ml = torch.zeros(B, T, dtype=torch.bool)
    ml[:, 50:] = True
    blk = util.load_sd(BiAttentionBlock(256, 256, 1024, 4), sd)
    rv, rl = restate.gdino_bi_attention(v, l, sd, "", 4, 1024, mask_v=mv, mask_l=ml)
    ov, ol = blk(v, l, attention_mask_v=mv, attention_mask_l=ml)
    util.assert_close(ov, rv, 1e-3, "explicit BiAttention, K-split text side: v")
    util.assert_close(ol, rl, 1e-3, "explicit BiAttention, K-split text side: l")


def case_gcp_bwd():
    import test_train_gpu as t
    t.test_gcp_block_backward_vs_autograd(CPU, 2, 64, 6)


def case_bert_bwd():
    from mqdet_b200.modeling.language_backbone.bert_backward import BertLayerTrain
    from mqdet_b200.modeling.language_backbone.modeling_bert_new import BertLayer
    from oracle import synth
    gen = synth.Gen(8)
    sd = synth.bert_layer_sd(gen, "")
    h, dy = gen.randn(2, 64, 768), gen.randn(2, 64, 768)
    am = torch.ones(2, 64)
    am[0, 50:] = 0
    hr = h.clone().requires_grad_(True)
    restate.bert_layer(hr, restate.extended_mask(am), sd, "", 12).backward(dy)
    tr = BertLayerTrain(util.load_sd(BertLayer(768, 12, 3072), sd))
    tr.forward(h, h.half(), am)
    util.assert_close(tr.backward(dy), hr.grad, 3e-3, "bert layer backward: dh")


def case_preselect_bwd():
    import test_train_gpu as t
    t.test_preselect_backward_vs_autograd(CPU, 2, 50, 237)


def case_lang_train():
    import test_train_gpu as t
    t.test_qvbert_encoder_backward_vs_autograd(CPU)


def case_ref_signatures():
    """The reference-signature entry points of the GroundingDINO modules (lists of [B,C,h,w] maps, sequence-first text tensors, boolean
    masks): Transformer.forward, TransformerEncoderLayer.forward, DeformableTransformerEncoderLayer.forward, MLP on fp32 input,
    GroundingDINO.forward(return_backbone_features=True)."""
    import test_gdino_asm_gpu as t
    from mqdet_b200.structures.image_list import ImageList
    from oracle import make_golden
    c = make_golden.case_inputs("gdino_transformer")
    sd, nq, el, dl = c["sd"], c["nq"], c["enc_layers"], c["dec_layers"]
    poss = [restate.position_embedding_sine_hw(m) for m in c["masks"]]
    T = t._build_transformer(sd, nq, el, dl, CPU)
    td = {"encoded_text": c["enc_text"].clone(), "text_token_mask": c["tmask"], "position_ids": c["pid"],
          "text_self_attention_masks": c["sam"]}
    hs, refs, hs_enc, ref_enc, init = T(c["srcs"], c["masks"], None, poss, None, None, td)
    ref = restate.gdino_transformer(c["srcs"], c["masks"], poss, c["enc_text"], c["tmask"], c["pid"], c["sam"], sd, num_queries=nq,
                                    enc_layers=el, dec_layers=dl)
    assert len(hs) == dl and len(refs) == dl + 1 and hs_enc.shape == (1, 2, nq, 256) and ref_enc.shape == (1, 2, nq, 4)
    util.assert_close(hs[-1], ref["hs"][-1], 2e-3, "Transformer.forward (reference signature): hs[-1]")
    util.assert_close(td["encoded_text"], ref["memory_text"], 2e-3, "Transformer.forward: text_dict['encoded_text'] replaced")
    pos = restate.sine_pos_embed(c["pid"][..., None].float(), 256, exchange_xy=False)
    o = T.encoder.text_layers[0](c["enc_text"].transpose(0, 1), src_mask=~c["sam"], src_key_padding_mask=~c["tmask"], pos=pos.transpose(0, 1))
    util.assert_close(o.transpose(0, 1), restate.gdino_text_enhancer_layer(c["enc_text"], pos, ~c["sam"], sd, "encoder.text_layers.0.", 4),
                      2e-3, "TransformerEncoderLayer.forward")
    shapes = [tuple(s.shape[-2:]) for s in c["srcs"]]
    srcf = torch.cat([s.flatten(2).transpose(1, 2) for s in c["srcs"]], 1)
    maskf = torch.cat([m.flatten(1) for m in c["masks"]], 1)
    posf = torch.cat([p.flatten(2).transpose(1, 2) for p in poss], 1)
    vr = torch.stack([restate.gdino_valid_ratio(m) for m in c["masks"]], 1)
    refp = restate.gdino_encoder_reference_points(shapes, vr)
    o = T.encoder.layers[0](srcf, posf, refp, torch.tensor(shapes), None, key_padding_mask=maskf)
    util.assert_close(o, restate.gdino_deformable_encoder_layer(srcf, posf, refp, shapes, maskf, sd, "encoder.layers.0."), 2e-3,
                      "DeformableTransformerEncoderLayer.forward")
    assert T.decoder.ref_point_head(torch.randn(3, 7, 512)).shape == (3, 7, 256)
    use_oracle_backbones()
    sdm, ids, am, pmap, bank, img = t._gdino_case(3, 1, 96, 128, 5, 1, 1, 30)
    model = t._build_model(sdm, 1, 1, 30, CPU)
    model.query_selector.set_query_bank(bank)
    res, maps = model(ImageList(img, [(96, 128)]), captions={"input_ids": ids, "attention_mask": am}, positive_map=pmap,
                      return_backbone_features=True)
    assert [tuple(m.shape) for m in maps] == [(1, 256, 12, 16), (1, 256, 6, 8), (1, 256, 3, 4), (1, 256, 2, 2)] and len(res) == 1


if __name__ == "__main__":
    name = sys.argv[1]
    globals()["case_" + name]()
    print("PASS", name)
