"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt by running the REFERENCE's own modules (unmodified, loaded
from /root/reference by oracle/ref_loader.py) on seeded synthetic weights/inputs (oracle/synth.py).

    python -m oracle.make_golden            # in the build container (needs /root/reference)

Each fixture stores the case description (seed, sizes) and a strided SUBSAMPLE of the reference output (full tensors
would be MBs); weights/inputs are regenerated from the seed by ``case_inputs`` below, which the tests import too.
"""
import math
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def ref_cfg():
    vq = types.SimpleNamespace(ENABLED=True, FIX_ATTN_GATE=-1.0, CONDITION_GATE=True, NONLINEAR_GATE=True, NO_CAT=True,
                               ADD_ADAPT_LAYER=False, RETURN_ATTN_GATE_VALUE=False, VISION_SCALE=1.0,
                               AUGMENT_IMAGE_WITH_QUERY=False, TEXT_DROPOUT=0.4, NEW_MASK_TOKEN=False, QUERY_FUSION=False,
                               SHARE_KV=False, NUM_QUERY_PER_CLASS=5)
    fuse = types.SimpleNamespace(STABLE_SOFTMAX_2D=False, CLAMP_MIN_FOR_UNDERFLOW=True, CLAMP_MAX_FOR_OVERFLOW=True,
                                 SEPARATE_BIDIRECTIONAL=False, DO_LANG_PROJ_OUTSIDE_CHECKPOINT=False)
    model = types.SimpleNamespace(DYHEAD=types.SimpleNamespace(FUSE_CONFIG=fuse, NUM_CONVS=6))
    return types.SimpleNamespace(VISION_QUERY=vq, MODEL=model)


def ref_head_cfg():
    """mq-glip-t configuration of the VL head (configs/pretrain/mq-glip-t.yaml + config/defaults.py) as an attribute
    namespace, including the keys only the reference's constructors read."""
    NS = types.SimpleNamespace
    fuse = NS(EARLY_FUSE_ON=True, TYPE="MHA-B", JOINT_EMB_SIZE=256, JOINT_EMB_DROPOUT=0.1, JOINT_OUT_SIZE=256,
              JOINT_MLP_LAYERS=2, USE_DOT_PRODUCT_TOKEN_LOSS=True, USE_FUSED_FEATURES_DOT_PRODUCT=True, USE_TOKEN_LOSS=False,
              USE_CONTRASTIVE_ALIGN_LOSS=False, USE_SHALLOW_CONTRASTIVE_LOSS=False, USE_BACKBONE_SHALLOW_CONTRASTIVE_LOSS=False,
              USE_CLASSIFICATION_LOSS=False, MLM_LOSS=False, MLM_LOSS_COEF=1.0, TOKEN_LOSS_WEIGHT=1.0,
              DOT_PRODUCT_TOKEN_LOSS_WEIGHT=1.0, SHALLOW_CONTRASTIVE_LOSS_WEIGHT=1.0, CONTRASTIVE_ALIGN_LOSS_WEIGHT=1.0,
              CONTRASTIVE_HIDDEN_DIM=64, ADD_LINEAR_LAYER=False, USE_LAYER_SCALE=True, STABLE_SOFTMAX_2D=False,
              CLAMP_MIN_FOR_UNDERFLOW=True, CLAMP_MAX_FOR_OVERFLOW=True, CLAMP_BERTATTN_MIN_FOR_UNDERFLOW=True,
              CLAMP_BERTATTN_MAX_FOR_OVERFLOW=True, CLAMP_DOT_PRODUCT=True, SEPARATE_BIDIRECTIONAL=False,
              DO_LANG_PROJ_OUTSIDE_CHECKPOINT=False)
    dyhead = NS(NUM_CLASSES=81, CHANNELS=256, NUM_CONVS=6, USE_GN=True, USE_SYNCBN=False, USE_NSYNCBN=False, USE_DYRELU=True,
                USE_DFCONV=True, USE_DYFUSE=True, USE_CHECKPOINT=False, CONV_FUNC="", TOPK=9, PRIOR_PROB=0.01, LOG_SCALE=0.0,
                SCORE_AGG="MEAN", FUSE_CONFIG=fuse)
    model = NS(DEVICE="cpu", RPN_ONLY=True, BACKBONE=NS(OUT_CHANNELS=256), GROUP_NORM=NS(NUM_GROUPS=16), DYHEAD=dyhead,
               LANGUAGE_BACKBONE=NS(LANG_DIM=768, MAX_QUERY_LEN=256, N_LAYERS=1, MODEL_TYPE="bert-base-uncased"),
               RPN=NS(ASPECT_RATIOS=(1.0,), SCALES_PER_OCTAVE=1, RETURN_FUSED_FEATURES=False),
               CLIP=NS(WIDTH=512, VOCAB_SIZE=49408))
    return NS(MODEL=model, VISION_QUERY=ref_cfg().VISION_QUERY)


def ref_detector_cfg():
    """ref_head_cfg() + the keys the detector, the anchor generator and the post-processor read (mq-glip-t.yaml)."""
    NS = types.SimpleNamespace
    cfg = ref_head_cfg()
    m = cfg.MODEL
    m.SWINT = NS(VERSION="v1")
    m.LANGUAGE_BACKBONE.PAD_MAX = True
    m.LANGUAGE_BACKBONE.MASK_SPECIAL = False
    m.LANGUAGE_BACKBONE.USE_CHECKPOINT = False
    m.RPN = NS(ASPECT_RATIOS=(1.0,), SCALES_PER_OCTAVE=1, RETURN_FUSED_FEATURES=False, ANCHOR_SIZES=(64, 128, 256, 512, 1024),
               ANCHOR_STRIDE=(8, 16, 32, 64, 128), STRADDLE_THRESH=0, OCTAVE=2.0, USE_FPN=True)
    m.ATSS = NS(INFERENCE_TH=0.05, INFERENCE_TH_TRAIN=0.0, PRE_NMS_TOP_N=1000, PRE_NMS_TOP_N_TRAIN=3000, NMS_TH=0.6,
                DETECTIONS_PER_IMG=100, POST_NMS_TOP_N_TRAIN=1000, NUM_CLASSES=81)
    m.ROI_MASK_HEAD = NS(PREDICTOR="MaskRCNNC4Predictor")
    cfg.TEST = NS(USE_MULTISCALE=False, MDETR_STYLE_AGGREGATE_CLASS_NUM=-1)
    cfg.GLIPKNOW = NS(PARALLEL_LANGUAGE_INPUT=False)
    cfg.DATASETS = NS(ONE_HOT=False)
    vq = cfg.VISION_QUERY
    for k, v in dict(QUERY_BANK_PATH="", LEARNABLE_BANK=False, ADD_VISION_LAYER=False, PURE_TEXT_RATE=0.0, RANDOM_KSHOT=False,
                     MASK_DURING_INFERENCE=False, GATE_REGULARIZATION=False, GATE_REGULARIZATION_SCALE=1.0).items():
        setattr(vq, k, v)
    return cfg


def canonical_detections(d):
    """[n, 6] (x1, y1, x2, y2, score, label) sorted by (score, label, x1): the reference's detection order depends on
    topk(sorted=False), so detections are compared as sets."""
    key = d[:, 4].double() * 1e6 + d[:, 5].double() * 1e-3 + d[:, 0].double() * 1e-9
    return d[torch.argsort(key)]


def dcn_stub(x, offset, mask, weight, bias, stride):
    """Stands in for the compiled modulated_deform_conv (pinned against the real kernel on the GPU,
    tests/test_ref_kernels_gpu.py): the flat per-image offset / (sigmoid-ed) mask buffers go to the oracle's restatement of
    the kernel, which indexes them with the OUTPUT strides exactly like deform_conv_kernel_cuda.cu:605-618 (so the DyConv[0]
    offset re-interpretation happens here too)."""
    from oracle import restate
    B = x.shape[0]
    return restate.dcn_v2(x, offset.reshape(B, -1), mask.reshape(B, -1), weight, bias, stride)


LEVELS_SMALL = [(20, 28), (10, 14), (5, 7), (3, 4), (2, 2)]  # a 160x224 "image" through strides 8..128


def sub(t, *steps):
    """strided subsample along the trailing dims"""
    idx = [slice(None)] * (t.dim() - len(steps)) + [slice(None, None, s) for s in steps]
    return t[tuple(idx)].contiguous().clone()


# ----------------------------------------------------------------------------------------------------------------------
# case inputs (shared with the tests)
# ----------------------------------------------------------------------------------------------------------------------
def case_inputs(name):
    if name == "gcp_block":
        gen = synth.Gen(1234)
        sd = synth.gcp_block_sd(gen)
        _, _, pmap = synth.prompt(10, 2, 256, gen)
        _, m = synth.vision_queries(pmap, 5, 256, 768, gen)
        B = 2
        mask = m.expand(B, -1, -1).clone()
        mask[0, 3] = 0
        mask[0, 7:10] = 0
        mask[1, 20:25] = 0
        vision = gen.randn(B, 50, 768)
        x = gen.randn(B, 256, 768)
        return dict(sd=sd, x=x, vision=vision, mask=mask)
    if name == "preselect":
        gen = synth.Gen(1235)
        sd = synth.preselect_sd(gen)
        return dict(sd=sd, vision=gen.randn(2, 50, 256, scale=0.5), image=gen.randn(2, 1117, 256))
    if name == "bi_attention":
        gen = synth.Gen(1236)
        sd = synth.bi_attention_sd(gen)
        B, T = 2, 256
        feats = [gen.randn(B, 256, h, w) for (h, w) in LEVELS_SMALL]
        l = gen.randn(B, T, 768)
        mask = torch.ones(B, T, dtype=torch.long)
        mask[0, 100:] = 0
        mask[1, 33:] = 0
        return dict(sd=sd, feats=feats, l=l, mask=mask)
    if name == "bert_layer":
        gen = synth.Gen(1237)
        sd = synth.bert_layer_sd(gen, "")
        B, T = 2, 256
        h = gen.randn(B, T, 768)
        am = torch.ones(B, T)
        am[0, 200:] = 0
        am[1, 33:] = 0
        return dict(sd=sd, h=h, am=am)
    if name == "swin_fpn":
        gen = synth.Gen(1238)
        sd = synth.swin_sd(gen)
        fsd = synth.fpn_sd(gen)
        img = gen.randn(2, 3, 150, 203, scale=1.0)  # not a multiple of 4/7/2: exercises every padding path
        return dict(sd=sd, fsd=fsd, img=img)
    if name == "dyconv":
        gen = synth.Gen(77)
        sd = synth.dyconv_sd(gen)
        return dict(sd=sd, feats=[gen.randn(2, 256, h, w) for h, w in LEVELS_SMALL])
    if name == "vldyhead":
        gen = synth.Gen(78)
        sd = synth.vldyhead_sd(gen, 6)
        B, T = 2, 256
        feats = [gen.randn(B, 256, h, w) for h, w in LEVELS_SMALL]
        hidden = gen.randn(B, T, 768)
        masks = torch.ones(B, T, dtype=torch.long)
        masks[0, 120:] = 0
        masks[1, 31:] = 0
        return dict(sd=sd, feats=feats, hidden=hidden, masks=masks)
    if name == "detector":
        gen = synth.Gen(71)
        sd = synth.detector_sd(gen, bias0=-1.0)  # enough (location, class) pairs above 0.05 to hit the top-k and the NMS
        ids, am, pmap = synth.prompt(12, 2, 256, gen)
        bank = synth.query_bank(pmap, 5, gen)
        return dict(sd=sd, ids=ids, am=am, pmap=pmap, bank=bank, img=synth.images(gen, 1, 160, 224), size=(160, 224))
    if name == "detector_bench":
        # BASELINE.json config 2 per image: ONE 800x1333 image (padded 800x1344), 80-class prompt, K = 5 queries per class,
        # head bias at the reference's PRIOR_PROB initialisation (what bench.py runs, at B = 8 copies of such images)
        gen = synth.Gen(72)
        sd = synth.detector_sd(gen, bias0=-math.log(99.0))
        ids, am, pmap = synth.prompt(80, 2, 256, gen)
        bank = synth.query_bank(pmap, 5, gen)
        return dict(sd=sd, ids=ids, am=am, pmap=pmap, bank=bank, img=synth.images(gen, 1, 800, 1333), size=(800, 1333))
    if name == "contrastive_embed":
        gen = synth.Gen(1239)
        B, Q, T, D = 2, 900, 195, 256
        mask = torch.ones(B, T, dtype=torch.bool)
        mask[0, 150:] = False
        mask[1, 17:] = False
        return dict(x=gen.randn(B, Q, D), y=gen.randn(B, T, D), mask=mask)
    if name == "gdino_transformer":
        # GroundingDINO Transformer.forward (BASELINE config 4): 2 encoder + 2 decoder layers, 4 small levels (image 1 padded on
        # the right / bottom), 20 queries, a prompt of 4-token categories with per-category masks and restarted position ids
        gen = synth.Gen(1240)
        nq, el, dl, B, Tt, used = 20, 2, 2, 2, 32, 25
        sd = synth.gdino_transformer_sd(gen, el, dl, nq=nq)
        shapes = ((12, 16), (6, 8), (3, 4), (2, 2))
        srcs = [gen.randn(B, 256, h, w) for h, w in shapes]
        masks = []
        for h, w in shapes:
            m = torch.zeros(B, h, w, dtype=torch.bool)
            m[1, :, int(w * 0.75):] = True
            m[1, int(h * 0.8):, :] = True
            masks.append(m)
        tmask = torch.ones(B, Tt, dtype=torch.bool)
        tmask[:, used:] = False
        pid = torch.zeros(B, Tt, dtype=torch.long)
        sam = torch.eye(Tt, dtype=torch.bool)[None].repeat(B, 1, 1)
        st = 1
        while st < used:
            e = min(st + 4, used)
            sam[:, st:e, st:e] = True
            pid[:, st:e] = torch.arange(e - st)
            st = e
        return dict(sd=sd, nq=nq, enc_layers=el, dec_layers=dl, srcs=srcs, masks=masks, enc_text=gen.randn(B, Tt, 256), tmask=tmask,
                    pid=pid, sam=sam)
    raise KeyError(name)


# ----------------------------------------------------------------------------------------------------------------------
# reference runs
# ----------------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def run_reference(name):
    from oracle import ref_loader as rl
    c = case_inputs(name)
    cfg = ref_cfg()
    if name == "gcp_block":
        m = rl.modeling_bert_new()
        blk = m.GatedCrossAttentionBlock(dim=768, cfg=cfg).eval()
        blk.load_state_dict(c["sd"], strict=True)
        y = blk(c["x"], c["vision"], c["mask"])
        s = blk.attn(c["x"], c["vision"], c["mask"])
        return dict(y=y, s=s)
    if name == "preselect":
        m = rl.modeling_bert_new()
        mod = m.PreSelectModule(dim=256, out_dim=768, cfg=cfg).eval()
        mod.load_state_dict(c["sd"], strict=True)
        return dict(vision=mod(c["vision"], c["image"])["vision"])
    if name == "bi_attention":
        fh = rl.fuse_helper()
        blk = fh.BiAttentionBlockForCheckpoint(v_dim=256, l_dim=768, embed_dim=2048, num_heads=8, hidden_dim=3072,
                                               dropout=0.1, drop_path=0.0, init_values=1.0 / 6, cfg=cfg).eval()
        blk.load_state_dict(c["sd"], strict=True)
        out = blk(*c["feats"], c["l"], c["mask"], None)
        v = torch.cat([o.flatten(2).transpose(1, 2) for o in out[:5]], dim=1)  # [B,N,256], P3->P7 row-major
        return dict(v=v, l=out[5])
    if name == "bert_layer":
        rb = rl.rpn_modeling_bert()
        from transformers import BertConfig
        config = BertConfig()
        att = rb.BertAttention(config, True, True).eval()
        inter = rb.BertIntermediate(config).eval()
        outp = rb.BertOutput(config).eval()
        sd = c["sd"]
        att.load_state_dict({k[len("attention."):]: v for k, v in sd.items() if k.startswith("attention.")}, strict=True)
        inter.load_state_dict({k[len("intermediate."):]: v for k, v in sd.items() if k.startswith("intermediate.")}, strict=True)
        outp.load_state_dict({k[len("output."):]: v for k, v in sd.items() if k.startswith("output.")}, strict=True)
        ext = (1.0 - c["am"][:, None, None, :]) * -10000.0
        # wiring of BertEncoderLayer.forward (maskrcnn_benchmark/modeling/rpn/vldyhead.py:264-301)
        a = att(c["h"], ext, None, output_attentions=False, past_key_value=None)[0]
        return dict(h=outp(inter(a), a))
    if name == "dyconv":
        vd = rl.vldyhead(dcn_stub)
        conv_func = lambda i, o, s: vd.Conv3x3Norm(i, o, s, deformable=True, bn_type=["gn", 16])  # noqa: E731
        mod = vd.DyConv(256, 256, conv_func=conv_func, use_dyrelu=True, use_dyfuse=True, use_deform=True).eval()
        mod.load_state_dict(c["sd"], strict=True)
        out = mod({"visual": [f.clone() for f in c["feats"]], "lang": None})["visual"]
        return dict(v=torch.cat([o.flatten(2).transpose(1, 2) for o in out], dim=1))
    if name == "vldyhead":
        import contextlib
        import io
        vd = rl.vldyhead(dcn_stub)
        with contextlib.redirect_stdout(io.StringIO()):  # the constructor prints "EARLY FUSION ON" per layer
            head = vd.VLDyHead(ref_head_cfg()).eval()
        head.load_state_dict(c["sd"], strict=True)
        lang = {"hidden": c["hidden"].clone(), "masks": c["masks"], "embedded": c["hidden"].clone()}
        out = head([f.clone() for f in c["feats"]], lang, embedding=lang["embedded"])
        flat = lambda xs: torch.cat([x.flatten(2).transpose(1, 2) for x in xs], dim=1)  # noqa: E731
        return dict(logits=torch.cat(out[6], dim=1), hidden=lang["hidden"], bbox=flat(out[1]), ctr=flat(out[2]),
                    per_level=(out[6], out[1], out[2]))
    if name in ("detector", "detector_bench"):
        import numpy as np
        from oracle import restate
        det = rl.detector(ref_detector_cfg(), dcn_stub, lambda b, s, l, t: restate.ml_nms(b, s, l, t), (c["ids"], c["am"]))
        full = dict(c["sd"])
        for k, v in det.state_dict().items():  # index / anchor buffers are not parameters
            if k not in full:
                assert k.endswith("relative_position_index") or "cell_anchors" in k, k
                full[k] = v
        det.load_state_dict(full, strict=True)
        det.query_selector.query_bank = {k: v.clone() for k, v in c["bank"].items()}
        np.random.seed(0)
        cap = {}
        # intermediates of the reference forward: language stream entering the head, head outputs (logits / fused hidden)
        h1 = det.language_backbone.register_forward_hook(lambda m, i, o: cap.__setitem__("lang_hidden", o["hidden"].clone()))
        h2 = det.rpn.head.register_forward_hook(lambda m, i, o: cap.__setitem__("head", o))
        h3 = det.rpn.head.register_forward_pre_hook(lambda m, i: cap.__setitem__("pyr", [f.clone() for f in i[0]]))
        il = sys.modules["maskrcnn_benchmark.structures.image_list"].ImageList(c["img"], [c["size"]])
        bl = det(il, captions=["a synthetic caption"], positive_map=c["pmap"])[0]
        for h in (h1, h2, h3):
            h.remove()
        d = torch.cat([bl.bbox, bl.get_field("scores")[:, None], bl.get_field("labels")[:, None].float()], 1)
        flat = lambda xs: torch.cat([x.flatten(2).transpose(1, 2) for x in xs], dim=1)  # noqa: E731
        return dict(det=canonical_detections(d), boxlist=bl, lang_hidden=cap["lang_hidden"], logits=torch.cat(cap["head"][6], 1),
                    pyramid=flat(cap["pyr"]), bbox=flat(cap["head"][1]), ctr=flat(cap["head"][2]))
    if name == "contrastive_embed":
        mod = rl.gdino_utils().ContrastiveEmbed(max_text_len=256)
        return dict(logits=mod(c["x"], {"encoded_text": c["y"], "text_token_mask": c["mask"]}))
    if name == "gdino_transformer":
        import warnings
        p = rl.gdino_package()
        kw = dict(d_model=256, nhead=8, dim_feedforward=2048, dropout=0.0, activation="relu", return_intermediate_dec=True, query_dim=4,
                  num_feature_levels=4, enc_n_points=4, dec_n_points=4, learnable_tgt_init=True, two_stage_type="standard",
                  embed_init_tgt=True, use_text_enhancer=True, use_fusion_layer=True, use_checkpoint=False, use_transformer_ckpt=False,
                  use_text_cross_attention=True, text_dropout=0.0, fusion_dropout=0.0, fusion_droppath=0.1)
        T = p.transformer.Transformer(num_queries=c["nq"], num_encoder_layers=c["enc_layers"], num_decoder_layers=c["dec_layers"], **kw).eval()
        be = p.utils.MLP(256, 256, 4, 3)  # shared by the decoder layers (groundingdino.py:247-252)
        T.decoder.bbox_embed = torch.nn.ModuleList([be for _ in range(c["dec_layers"])])
        T.decoder.class_embed = torch.nn.ModuleList([p.utils.ContrastiveEmbed() for _ in range(c["dec_layers"])])
        T.enc_out_bbox_embed = p.utils.MLP(256, 256, 4, 3)
        T.enc_out_class_embed = p.utils.ContrastiveEmbed()
        T.load_state_dict({k: c["sd"][k] for k in T.state_dict()}, strict=True)
        import importlib
        pe = importlib.import_module("ref_gdino_pkg.backbone.position_encoding").PositionEmbeddingSineHW(128, 20, 20, normalize=True)
        misc = importlib.import_module("groundingdino_new.util.misc")
        poss = [pe(misc.NestedTensor(s, m)) for s, m in zip(c["srcs"], c["masks"])]
        td = {"encoded_text": c["enc_text"].clone(), "text_token_mask": c["tmask"], "position_ids": c["pid"],
              "text_self_attention_masks": c["sam"]}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            hs, refs, _, _, _ = T(c["srcs"], c["masks"], None, poss, None, None, td)
        return dict(hs_last=hs[-1], ref_last=refs[-1], text=td["encoded_text"])
    if name == "swin_fpn":
        sw = rl.swint()
        body = sw.SwinTransformer(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7,
                                  drop_path_rate=0.0, frozen_stages=-1, use_checkpoint=False)
        body.eval()  # the reference's train() override returns None, so no chaining
        sdb = dict(c["sd"])
        for k, v in body.state_dict().items():
            if k.endswith("relative_position_index"):
                sdb[k] = v
        body.load_state_dict(sdb, strict=True)
        outs = body(c["img"])
        fpn_mod = rl.fpn()
        conv_block = lambda i, o, k, s=1: torch.nn.Conv2d(i, o, k, s, padding=(k - 1) // 2)  # noqa: E731
        f = fpn_mod.FPN([0, 192, 384, 768], 256, conv_block, top_blocks=fpn_mod.LastLevelP6P7(256, 256)).eval()
        f.load_state_dict(c["fsd"], strict=True)
        pyr = f(outs)
        res = {f"c{i + 2}": o for i, o in enumerate(outs)}
        res.update({f"p{i + 3}": o for i, o in enumerate(pyr)})
        return res
    raise KeyError(name)


SUBSAMPLE = {"gcp_block": {"y": (4, 8), "s": (4, 8)}, "preselect": {"vision": (1, 8)},
             "bi_attention": {"v": (3, 4), "l": (4, 8)}, "bert_layer": {"h": (4, 8)},
             "contrastive_embed": {"logits": (9, 1)},
             "gdino_transformer": {"hs_last": (1, 1), "ref_last": (1, 1), "text": (1, 1)},
             "dyconv": {"v": (3, 4)},
             "detector": {"det": (1, 1)},
             "detector_bench": {"det": (1, 1), "lang_hidden": (2, 8), "logits": (32, 2), "pyramid": (32, 4), "bbox": (16, 1),
                                "ctr": (16, 1)},
             "vldyhead": {"logits": (3, 4), "hidden": (4, 8), "bbox": (3, 1), "ctr": (3, 1)},
             "swin_fpn": {"c3": (4, 2, 2), "c4": (4, 1, 1), "c5": (8, 1, 1), "p3": (4, 2, 2), "p4": (4, 1, 1), "p5": (4, 1, 1),
                          "p6": (2, 1, 1), "p7": (1, 1, 1)}}


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = sys.argv[1:]
    for name, subs in SUBSAMPLE.items():
        if only and name not in only:
            continue
        out = run_reference(name)
        fx = {"case": name, "subsample": subs, "torch": str(torch.__version__)}
        for k, steps in subs.items():
            fx[k] = sub(out[k].float(), *steps)
            fx[k + "_absmax"] = out[k][torch.isfinite(out[k])].abs().max().item()
        path = os.path.join(GOLDEN_DIR, f"{name}.pt")
        torch.save(fx, path)
        print(name, {k: tuple(fx[k].shape) for k in subs}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
