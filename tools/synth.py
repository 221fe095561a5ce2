"""Deterministic synthetic weights and inputs for the tests, the oracle scripts and bench.py.

No checkpoints, tokenizer vocabularies or datasets exist offline, so every test/bench input is generated from a
seeded ``torch.Generator`` on the CPU.  Weight dicts use the reference's ``state_dict`` key names so they load into
the unmodified reference modules (pinning) as well as into mqdet_b200's drop-in modules.
Gates are initialised NON-zero (the reference initialises them to 0, which would make GCP parity vacuous:
modeling_bert_new.py:274,289).
"""
import math

import torch


class Gen:
    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)

    def linear(self, out_f, in_f, bias=False, sd=None, name=None, bias_scale=0.02):
        bound = 1.0 / math.sqrt(in_f)
        w = (torch.rand(out_f, in_f, generator=self.g) * 2 - 1) * bound
        sd[name + ".weight"] = w
        if bias:
            sd[name + ".bias"] = torch.randn(out_f, generator=self.g) * bias_scale

    def xavier(self, out_f, in_f, sd, name, bias_scale=0.02):
        bound = math.sqrt(6.0 / (in_f + out_f))
        sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=self.g) * 2 - 1) * bound
        sd[name + ".bias"] = torch.randn(out_f, generator=self.g) * bias_scale

    def norm(self, dim, sd, name):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(dim, generator=self.g)
        sd[name + ".bias"] = 0.1 * torch.randn(dim, generator=self.g)

    def randn(self, *shape, scale=1.0):
        return torch.randn(*shape, generator=self.g) * scale


def gcp_block_sd(gen, p="", dim=768, heads=8, dim_head=64, ff_mult=4, sd=None):
    """GatedCrossAttentionBlock parameters (modeling_bert_new.py:256-296): 6 592 897 params at dim=768."""
    sd = {} if sd is None else sd
    inner = heads * dim_head
    gen.norm(dim, sd, p + "attn.norm")
    gen.norm(dim, sd, p + "attn.norm_kv")
    gen.linear(inner, dim, sd=sd, name=p + "attn.to_q")
    gen.linear(2 * inner, dim, sd=sd, name=p + "attn.to_kv")
    gen.linear(dim, inner, sd=sd, name=p + "attn.to_out")
    gen.norm(dim, sd, p + "attn_gate.norm")
    gen.linear(dim // 2, dim, sd=sd, name=p + "attn_gate.linear1")
    sd[p + "attn_gate.linear2.weight"] = gen.randn(1, dim // 2, scale=0.1)
    gen.norm(dim, sd, p + "ff.norm")
    gen.linear(dim * ff_mult, dim, sd=sd, name=p + "ff.linear1")
    gen.linear(dim, dim * ff_mult, sd=sd, name=p + "ff.linear2")
    sd[p + "ff_gate"] = torch.tensor([0.3])
    return sd


def preselect_sd(gen, p="", dim=256, out_dim=768, heads=8, dim_head=32, ff_mult=4, num_layers=2, sd=None):
    """PreSelectModule parameters (modeling_bert_new.py:377-431)."""
    sd = {} if sd is None else sd
    inner = heads * dim_head
    for i in range(num_layers):
        od = out_dim if i == num_layers - 1 else dim
        lp = f"{p}layers.{i}."
        gen.norm(dim, sd, lp + "image_condition.norm")
        gen.norm(dim, sd, lp + "image_condition.norm_kv")
        gen.linear(inner, dim, sd=sd, name=lp + "image_condition.to_q")
        gen.linear(2 * inner, dim, sd=sd, name=lp + "image_condition.to_kv")
        gen.linear(od, inner, sd=sd, name=lp + "image_condition.to_out")
        gen.norm(od, sd, lp + "ff.norm")
        gen.linear(od * ff_mult, od, sd=sd, name=lp + "ff.linear1")
        gen.linear(od, od * ff_mult, sd=sd, name=lp + "ff.linear2")
        if od != dim:
            gen.linear(od, dim, sd=sd, name=lp + "res_mapping")
    return sd


def bert_layer_sd(gen, p, dim=768, inter=3072, sd=None, std=0.02):
    """One BERT layer, HF key names; normal(0, 0.02) like BertPreTrainedModel._init_weights, non-trivial LN/bias."""
    sd = {} if sd is None else sd
    for name, (o, i) in {"attention.self.query": (dim, dim), "attention.self.key": (dim, dim),
                         "attention.self.value": (dim, dim), "attention.output.dense": (dim, dim),
                         "intermediate.dense": (inter, dim), "output.dense": (dim, inter)}.items():
        # 0.02 std gives near-uniform attention; use a livelier scale so softmax parity is meaningful
        sd[p + name + ".weight"] = gen.randn(o, i, scale=std * 2.5)
        sd[p + name + ".bias"] = gen.randn(o, scale=0.02)
    gen.norm(dim, sd, p + "attention.output.LayerNorm")
    gen.norm(dim, sd, p + "output.LayerNorm")
    return sd


def qvbert_sd(gen, vocab=30522, dim=768, layers=12, max_pos=512, start_qv=6):
    """QVBertModel parameters (modeling_bert_new.py:642-660): BERT-base + 6 GCP blocks + PreSelect."""
    sd = {}
    sd["embeddings.word_embeddings.weight"] = gen.randn(vocab, dim, scale=0.5)
    sd["embeddings.position_embeddings.weight"] = gen.randn(max_pos, dim, scale=0.05)
    sd["embeddings.token_type_embeddings.weight"] = gen.randn(2, dim, scale=0.05)
    gen.norm(dim, sd, "embeddings.LayerNorm")
    for i in range(layers):
        bert_layer_sd(gen, f"encoder.layer.{i}.", dim, 4 * dim, sd)
    for i in range(layers - start_qv):
        gcp_block_sd(gen, f"encoder.qv_layer.{i}.", dim, sd=sd)
    preselect_sd(gen, "pre_select.", 256, dim, sd=sd)
    return sd


def bi_attention_sd(gen, p="", v_dim=256, l_dim=768, embed=2048, num_convs=6, sd=None):
    """BiAttentionBlockForCheckpoint parameters (fuse_helper.py:345-375, xavier init :207-216, gamma=1/NUM_CONVS)."""
    sd = {} if sd is None else sd
    gen.norm(v_dim, sd, p + "layer_norm_v")
    gen.norm(l_dim, sd, p + "layer_norm_l")
    gen.xavier(embed, v_dim, sd, p + "attn.v_proj")
    gen.xavier(embed, l_dim, sd, p + "attn.l_proj")
    gen.xavier(embed, v_dim, sd, p + "attn.values_v_proj")
    gen.xavier(embed, l_dim, sd, p + "attn.values_l_proj")
    gen.xavier(v_dim, embed, sd, p + "attn.out_v_proj")
    gen.xavier(l_dim, embed, sd, p + "attn.out_l_proj")
    sd[p + "gamma_v"] = torch.full((v_dim,), 1.0 / num_convs) * (1 + 0.1 * gen.randn(v_dim))
    sd[p + "gamma_l"] = torch.full((l_dim,), 1.0 / num_convs) * (1 + 0.1 * gen.randn(l_dim))
    return sd


def dot_head_sd(gen, p="", l_dim=768, channels=256, sd=None):
    """dot_product_projection_text, bias_lang, bias0, log_scale (vldyhead.py:711-720)."""
    sd = {} if sd is None else sd
    gen.linear(channels, l_dim, bias=True, sd=sd, name=p + "dot_product_projection_text")
    sd[p + "dot_product_projection_text.weight"] *= 20.0  # logits spread over several units so scores/top-k/NMS are non-degenerate
    sd[p + "bias_lang"] = gen.randn(l_dim, scale=0.3)
    sd[p + "bias0"] = torch.tensor([-math.log((1 - 0.01) / 0.01)])
    sd[p + "log_scale"] = torch.tensor([0.0])
    return sd


# ------------------------------------------------------------------------------------------------------------------
# prompts / queries
# ------------------------------------------------------------------------------------------------------------------
def prompt(num_classes, tokens_per_class=2, T=256, gen=None, vocab_lo=1996, vocab_hi=29000):
    """Synthetic tokenised caption "c1. c2. ..." (SURVEY.md §8d): [CLS]=101, class c -> tokens 1+3c..,
    '.'=1012 separators, [SEP]=102, [PAD]=0.  Returns (input_ids [1,T], attention_mask [1,T], positive_map)."""
    ids = torch.zeros(1, T, dtype=torch.long)
    ids[0, 0] = 101
    pos = 1
    positive_map = {}
    for c in range(num_classes):
        toks = list(range(pos, pos + tokens_per_class))
        positive_map[c + 1] = toks
        ids[0, toks] = torch.randint(vocab_lo, vocab_hi, (tokens_per_class,), generator=gen.g)
        pos += tokens_per_class
        ids[0, pos] = 1012
        pos += 1
    assert pos < T
    ids[0, pos] = 102
    mask = torch.zeros(1, T, dtype=torch.long)
    mask[0, : pos + 1] = 1
    return ids, mask, positive_map


def chunked_prompts(num_classes, chunk, T=256, gen=None, vocab_lo=1996, vocab_hi=29000):
    """LVIS-style chunked vocabulary (engine/inference.py:165-283: TEST.CHUNKED_EVALUATION classes per prompt): classes
    1..num_classes in chunks of ``chunk``; class c has 2 + (c % 2) tokens, so token positions differ between chunks.
    Returns [(input_ids [1,T], attention_mask [1,T], positive_map {global label: [token positions]}), ...]."""
    out = []
    for c0 in range(0, num_classes, chunk):
        ids = torch.zeros(1, T, dtype=torch.long)
        ids[0, 0] = 101
        pos = 1
        pm = {}
        for c in range(c0 + 1, min(c0 + chunk, num_classes) + 1):
            n = 2 + (c % 2)
            toks = list(range(pos, pos + n))
            pm[c] = toks
            ids[0, toks] = torch.randint(vocab_lo, vocab_hi, (n,), generator=gen.g)
            pos += n
            ids[0, pos] = 1012
            pos += 1
        assert pos < T
        ids[0, pos] = 102
        mask = torch.zeros(1, T, dtype=torch.long)
        mask[0, : pos + 1] = 1
        out.append((ids, mask, pm))
    return out


def vision_queries(positive_map, K, T=256, dim=256, gen=None):
    """What QuerySelector.forward returns for one image (query_selector.py:40-116): queries [1, V, dim] and the
    binarised mask [1, V, T] with 1 on the token positions of the query's class."""
    V = len(positive_map) * K
    q = gen.randn(1, V, dim, scale=0.5)
    m = torch.zeros(1, V, T)
    for ci, (label, toks) in enumerate(sorted(positive_map.items())):
        m[0, ci * K:(ci + 1) * K, toks] = 1.0
    return q, m


def dyconv_sd(gen, p="", C=256, sd=None):
    """DyConv parameters (vldyhead.py:155-204; DYReLU layers/dyrelu.py:38-78).  Livelier than the reference's
    std-0.01 init so that offsets move by whole pixels, masks vary and the DyReLU/attention branches are exercised."""
    sd = {} if sd is None else sd
    for k in range(3):
        sd[f"{p}DyConv.{k}.conv.weight"] = gen.randn(C, C, 3, 3, scale=0.03)
        sd[f"{p}DyConv.{k}.conv.bias"] = gen.randn(C, scale=0.05)
        gen.norm(C, sd, f"{p}DyConv.{k}.bn")
    sd[p + "AttnConv.1.weight"] = gen.randn(1, C, 1, 1, scale=0.3)
    sd[p + "AttnConv.1.bias"] = gen.randn(1, scale=0.5)
    sd[p + "relu.fc.0.weight"] = gen.randn(C // 4, C, scale=0.2)
    sd[p + "relu.fc.0.bias"] = gen.randn(C // 4, scale=0.1)
    sd[p + "relu.fc.2.weight"] = gen.randn(4 * C, C // 4, scale=0.3)
    sd[p + "relu.fc.2.bias"] = gen.randn(4 * C, scale=0.3)
    sd[p + "offset.weight"] = gen.randn(27, C, 3, 3, scale=0.02)
    sd[p + "offset.bias"] = gen.randn(27, scale=0.5)
    return sd


def vldyhead_sd(gen, num_convs=6, C=256, l_dim=768, num_classes=80):
    """VLDyHead parameters (vldyhead.py:594-767) with the reference's key names."""
    sd = {}
    for i in range(num_convs):
        bi_attention_sd(gen, f"dyhead_tower.{3 * i}.b_attn.", C, l_dim, 2048, num_convs, sd)
        bert_layer_sd(gen, f"dyhead_tower.{3 * i + 1}.", l_dim, 4 * l_dim, sd)
        dyconv_sd(gen, f"dyhead_tower.{3 * i + 2}.", C, sd)
    sd["cls_logits.weight"] = gen.randn(num_classes, C, 1, 1, scale=0.01)
    sd["cls_logits.bias"] = torch.full((num_classes,), -math.log(99.0))
    sd["bbox_pred.weight"] = gen.randn(4, C, 1, 1, scale=0.05)
    sd["bbox_pred.bias"] = gen.randn(4, scale=0.1)
    sd["centerness.weight"] = gen.randn(1, C, 1, 1, scale=0.05)
    sd["centerness.bias"] = gen.randn(1, scale=0.1)
    dot_head_sd(gen, "", l_dim, C, sd)
    for l in range(5):
        sd[f"scales.{l}.scale"] = torch.tensor([1.0 + 0.1 * l])
    return sd


def msda_sd(gen, p="", embed=256, heads=8, levels=4, points=4, sd=None):
    """MultiScaleDeformableAttention parameters (ms_deform_attn.py:186-220), livelier than the reference's zero-initialised
    offset / weight projections so that sampling locations and attention weights depend on the query."""
    sd = {} if sd is None else sd
    n = heads * levels * points
    sd[p + "sampling_offsets.weight"] = gen.randn(n * 2, embed, scale=0.05)
    sd[p + "sampling_offsets.bias"] = gen.randn(n * 2, scale=1.5)
    sd[p + "attention_weights.weight"] = gen.randn(n, embed, scale=0.08)
    sd[p + "attention_weights.bias"] = gen.randn(n, scale=0.3)
    gen.xavier(embed, embed, sd, p + "value_proj")
    gen.xavier(embed, embed, sd, p + "output_proj")
    return sd


def swin_sd(gen, depths=(2, 2, 6, 2), heads=(3, 6, 12, 24), embed=96, ws=7, p=""):
    """SwinTransformer parameters with the reference's key names (swint.py), livelier than trunc_normal(0.02)."""
    sd = {}
    sd[p + "patch_embed.proj.weight"] = gen.randn(embed, 3, 4, 4, scale=0.15)
    sd[p + "patch_embed.proj.bias"] = gen.randn(embed, scale=0.05)
    gen.norm(embed, sd, p + "patch_embed.norm")
    for i, depth in enumerate(depths):
        C = embed * 2 ** i
        for j in range(depth):
            bp = f"{p}layers.{i}.blocks.{j}."
            gen.norm(C, sd, bp + "norm1")
            sd[bp + "attn.relative_position_bias_table"] = gen.randn((2 * ws - 1) ** 2, heads[i], scale=0.5)
            sd[bp + "attn.qkv.weight"] = gen.randn(3 * C, C, scale=1.5 / math.sqrt(C))
            sd[bp + "attn.qkv.bias"] = gen.randn(3 * C, scale=0.2)
            sd[bp + "attn.proj.weight"] = gen.randn(C, C, scale=1.0 / math.sqrt(C))
            sd[bp + "attn.proj.bias"] = gen.randn(C, scale=0.05)
            gen.norm(C, sd, bp + "norm2")
            sd[bp + "mlp.fc1.weight"] = gen.randn(4 * C, C, scale=1.0 / math.sqrt(C))
            sd[bp + "mlp.fc1.bias"] = gen.randn(4 * C, scale=0.05)
            sd[bp + "mlp.fc2.weight"] = gen.randn(C, 4 * C, scale=0.5 / math.sqrt(4 * C))
            sd[bp + "mlp.fc2.bias"] = gen.randn(C, scale=0.05)
        if i < len(depths) - 1:
            gen.norm(4 * C, sd, f"{p}layers.{i}.downsample.norm")
            sd[f"{p}layers.{i}.downsample.reduction.weight"] = gen.randn(2 * C, 4 * C, scale=1.0 / math.sqrt(4 * C))
        if i > 0:  # norm0 is nn.Identity for *-RETINANET backbones (swint.py:547-548)
            gen.norm(C, sd, f"{p}norm{i}")
    return sd


def fpn_sd(gen, in_channels=(192, 384, 768), C=256, p=""):
    """FPN + LastLevelP6P7 parameters (fpn.py), kaiming-uniform-like scale with non-zero biases."""
    sd = {}
    for k, cin in zip((2, 3, 4), in_channels):
        sd[f"{p}fpn_inner{k}.weight"] = gen.randn(C, cin, 1, 1, scale=1.0 / math.sqrt(cin))
        sd[f"{p}fpn_inner{k}.bias"] = gen.randn(C, scale=0.05)
        sd[f"{p}fpn_layer{k}.weight"] = gen.randn(C, C, 3, 3, scale=1.0 / math.sqrt(9 * C))
        sd[f"{p}fpn_layer{k}.bias"] = gen.randn(C, scale=0.05)
    for k in ("p6", "p7"):
        sd[f"{p}top_blocks.{k}.weight"] = gen.randn(C, C, 3, 3, scale=1.0 / math.sqrt(9 * C))
        sd[f"{p}top_blocks.{k}.bias"] = gen.randn(C, scale=0.05)
    return sd


def detector_sd(gen, num_convs=6, bias0=None, swin=None):
    """Full MQ-GLIP-T parameter set with the reference's key names (277 M parameters).  ``swin`` = dict(depths, heads, embed,
    ws) selects another backbone (MQ-GLIP-L: depths (2,2,18,2), heads (6,12,24,48), embed 192, ws 12; with num_convs=8)."""
    sd = {}
    swin = swin or {}
    embed = swin.get("embed", 96)
    sd.update({"backbone.body." + k: v for k, v in swin_sd(gen, **swin).items()})
    sd.update({"backbone.fpn." + k: v for k, v in fpn_sd(gen, in_channels=(2 * embed, 4 * embed, 8 * embed)).items()})
    sd.update({"language_backbone.body.model." + k: v for k, v in qvbert_sd(gen).items()})
    head = vldyhead_sd(gen, num_convs)
    if bias0 is not None:
        head["bias0"] = torch.tensor([float(bias0)])
    sd.update({"rpn.head." + k: v for k, v in head.items()})
    return sd


def query_bank(positive_map, K, gen, dim=256):
    """{label: FloatTensor[K, 1, dim]} — the on-disk bank format of extract_vision_query (query_selector.py:24,37)."""
    return {label: gen.randn(K, 1, dim, scale=0.5) for label in sorted(positive_map)}


def images(gen, B, h, w, size_divisibility=32, mean=(103.530, 116.280, 123.675), std=(57.375, 57.120, 58.395)):
    """Synthetic BGR-255 images normalised like the reference transform, zero-padded to a multiple of 32
    (configs/pretrain/mq-glip-t.yaml:84-85,95)."""
    raw = torch.rand(B, 3, h, w, generator=gen.g) * 255.0
    x = (raw - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    H = -(-h // size_divisibility) * size_divisibility
    W = -(-w // size_divisibility) * size_divisibility
    out = torch.zeros(B, 3, H, W)
    out[:, :, :h, :w] = x
    return out


# ------------------------------------------------------------------------------------------------------------------
# MQ-GroundingDINO-T (BASELINE config 4)
# ------------------------------------------------------------------------------------------------------------------
def mha_sd(gen, p, embed=256, sd=None):
    """torch.nn.MultiheadAttention parameters (in_proj_weight / in_proj_bias / out_proj)."""
    sd[p + "in_proj_weight"] = gen.randn(3 * embed, embed, scale=1.0 / math.sqrt(embed))
    sd[p + "in_proj_bias"] = gen.randn(3 * embed, scale=0.05)
    gen.xavier(embed, embed, sd, p + "out_proj")
    return sd


def mlp_sd(gen, p, dims, sd, last_scale=None):
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        gen.xavier(b, a, sd, f"{p}layers.{i}")
    if last_scale is not None:
        i = len(dims) - 2
        sd[f"{p}layers.{i}.weight"] = gen.randn(dims[-1], dims[-2], scale=last_scale)
        sd[f"{p}layers.{i}.bias"] = gen.randn(dims[-1], scale=last_scale)
    return sd


def gdino_transformer_sd(gen, enc_layers=6, dec_layers=6, E=256, ffn=2048, nq=900, levels=4, sd=None, p=""):
    """Transformer parameters with the reference's key names (transformer.py), plus ``enc_out_bbox_embed`` and the SHARED
    ``decoder.bbox_embed.{i}`` as GroundingDINO registers them (groundingdino.py:247-275).  Gammas, sampling offsets, attention
    weights and the box heads are livelier than the reference's (near-)zero initialisation so parity is not vacuous."""
    sd = {} if sd is None else sd
    sd[p + "level_embed"] = gen.randn(levels, E, scale=0.5)
    sd[p + "tgt_embed.weight"] = gen.randn(nq, E)
    gen.xavier(E, E, sd, p + "enc_output")
    gen.norm(E, sd, p + "enc_output_norm")
    for i in range(enc_layers):
        q = f"{p}encoder.layers.{i}."
        msda_sd(gen, q + "self_attn.", E, 8, levels, 4, sd)
        gen.norm(E, sd, q + "norm1")
        gen.xavier(ffn, E, sd, q + "linear1")
        gen.xavier(E, ffn, sd, q + "linear2")
        gen.norm(E, sd, q + "norm2")
        q = f"{p}encoder.text_layers.{i}."
        mha_sd(gen, q + "self_attn.", E, sd)
        gen.xavier(ffn // 2, E, sd, q + "linear1")
        gen.xavier(E, ffn // 2, sd, q + "linear2")
        gen.norm(E, sd, q + "norm1")
        gen.norm(E, sd, q + "norm2")
        q = f"{p}encoder.fusion_layers.{i}."
        bi_attention_sd(gen, q, E, E, ffn // 2, enc_layers, sd)
    for i in range(dec_layers):
        q = f"{p}decoder.layers.{i}."
        msda_sd(gen, q + "cross_attn.", E, 8, levels, 4, sd)
        gen.norm(E, sd, q + "norm1")
        mha_sd(gen, q + "ca_text.", E, sd)
        gen.norm(E, sd, q + "catext_norm")
        mha_sd(gen, q + "self_attn.", E, sd)
        gen.norm(E, sd, q + "norm2")
        gen.xavier(ffn, E, sd, q + "linear1")
        gen.xavier(E, ffn, sd, q + "linear2")
        gen.norm(E, sd, q + "norm3")
    gen.norm(E, sd, p + "decoder.norm")
    mlp_sd(gen, p + "decoder.ref_point_head.", (2 * E, E, E), sd)
    mlp_sd(gen, p + "enc_out_bbox_embed.", (E, E, E, 4), sd, last_scale=0.02)
    shared = mlp_sd(gen, "", (E, E, E, 4), {}, last_scale=0.02)
    for i in range(dec_layers):
        for k, v in shared.items():
            sd[f"{p}decoder.bbox_embed.{i}.{k}"] = v
    return sd


def gdino_sd(gen, enc_layers=6, dec_layers=6, nq=900):
    """Full MQ-GroundingDINO-T parameter set with the reference's key names (groundingdino.py): Swin-T backbone (3 outputs),
    input projections + GroupNorm, BertModelWarper(QVBertModel) incl. the unused pooler, feat_map, transformer, box heads."""
    sd = {"backbone.0." + k: v for k, v in swin_sd(gen).items()}
    for l, c in enumerate((192, 384, 768)):
        sd[f"input_proj.{l}.0.weight"] = gen.randn(256, c, 1, 1, scale=1.0 / math.sqrt(c))
        sd[f"input_proj.{l}.0.bias"] = gen.randn(256, scale=0.05)
        gen.norm(256, sd, f"input_proj.{l}.1")
    sd["input_proj.3.0.weight"] = gen.randn(256, 768, 3, 3, scale=1.0 / math.sqrt(9 * 768))
    sd["input_proj.3.0.bias"] = gen.randn(256, scale=0.05)
    gen.norm(256, sd, "input_proj.3.1")
    sd.update({"bert." + k: v for k, v in qvbert_sd(gen).items()})
    gen.xavier(768, 768, sd, "bert.pooler.dense")
    gen.xavier(256, 768, sd, "feat_map")
    gdino_transformer_sd(gen, enc_layers, dec_layers, nq=nq, sd=sd, p="transformer.")
    for i in range(dec_layers):
        for j in range(3):
            for t in ("weight", "bias"):
                sd[f"bbox_embed.{i}.layers.{j}.{t}"] = sd[f"transformer.decoder.bbox_embed.{i}.layers.{j}.{t}"]
    return sd


def rgb_images(gen, B, h, w, size_divisibility=32):
    """Synthetic RGB [0,1] images normalised like configs/pretrain/mq-groundingdino-t.yaml:45-46, zero-padded to a multiple of 32."""
    return images(gen, B, h, w, size_divisibility, mean=(0.485 * 255, 0.456 * 255, 0.406 * 255), std=(0.229 * 255, 0.224 * 255, 0.225 * 255))
