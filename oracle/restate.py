"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the MQ-Det hot path (the parity oracle).

Every function restates, in plain PyTorch-on-CPU tensor algebra, what the cited reference code computes; weights are
passed as a flat ``{name: tensor}`` dict whose keys are the reference's own ``state_dict`` names (so the same dict can
be loaded into the unmodified reference module, tests/test_oracle_pinning.py).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import this module; the
product path (mqdet_b200/) never does.

Pinning status (the reference ships no tests or golden vectors) — every stage is checked against the
reference's OWN code executed in the build container (tests/test_oracle_pinning.py, tests/test_host_cpu.py; loaders in
oracle/ref_loader.py) and through fixtures recorded from it (tests/golden/*.pt by oracle/make_golden.py):
  QuerySelector, PreSelect, GCP block + sparse attention + index table, QVBertEncoder / QVBertModel.forward, BiAttention,
  BertEncoderLayer wiring, DyConv.forward, the whole VLDyHead.forward, ATSSPostProcessor.forward (+ BoxCoder, kthvalue cut),
  AnchorGenerator, Swin-T + FPN (bit-identical), GroundingDINO ContrastiveEmbed, BoxList / to_image_list, and the whole
  GeneralizedVLRCNN_New.forward.
Substitutions needed to run them on CPU here: the compiled ml_nms / DCNv2 kernels (-> this file's restatements, which are
pinned on the GPU against the reference's CUDA sources built by oracle/build_ref.py, tests/test_ref_kernels_gpu.py) and
transformers-4 API differences (BertLayer positional signature -> the reference's in-repo copy of the layer,
get_extended_attention_mask, BertConfig.from_pretrained offline, tokenizer).  The composition itself (`detector` below) is
pinned against one whole eval forward of the reference's GeneralizedVLRCNN_New assembled from its own parts
(test_detector_vs_reference; fixture tests/golden/detector.pt).  "parity unpinned": nothing on the SURVEY §8 path.
Paths are relative to the MQ-Det repository root.
"""
import math

import torch
import torch.nn.functional as F


def _ln(x, sd, name, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _lin(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


# ------------------------------------------------------------------------------------------------------------------
# GCP  — maskrcnn_benchmark/modeling/language_backbone/modeling_bert_new.py
# ------------------------------------------------------------------------------------------------------------------
def gcp_index(mask):
    """get_index_with_padding_batch (:40-63) on mask [B, V, T]: per text token the ascending indices of the vision
    queries whose mask is non-zero, padded with V up to S = the global max count.  Returns int64 [B, T, S]."""
    B, V, T = mask.shape
    nz = mask.transpose(1, 2) != 0  # [B, T, V]
    S = int(nz.sum(-1).max())
    idx = torch.full((B, T, max(S, 0)), V, dtype=torch.long)
    for b in range(B):
        for t in range(T):
            ids = torch.nonzero(nz[b, t]).flatten()
            idx[b, t, : ids.numel()] = ids
    return idx


def gcp_sparse_attention(x, vision, mask, sd, p="", heads=8, dim_head=64):
    """MaskedCrossAttention.forward with spase_forward=True (:162-248).  x [B,T,D], vision [B,V,D], mask [B,V,T]."""
    B, T, D = x.shape
    V = vision.shape[1]
    idx = gcp_index(mask)  # [B,T,S]
    S = idx.shape[2]
    vis_pad = torch.cat([vision, vision.new_zeros(B, 1, D)], dim=1)  # zero padding slot BEFORE norm_kv (:176-180)
    xn = _ln(x, sd, p + "norm")
    kn = _ln(vis_pad, sd, p + "norm_kv")  # unique rows only: K/V depend on the query, not on the token
    q = F.linear(xn, sd[p + "to_q.weight"]) * (dim_head ** -0.5)
    kv = F.linear(kn, sd[p + "to_kv.weight"])
    inner = heads * dim_head
    k_all, v_all = kv[..., :inner], kv[..., inner:]
    out = x.new_zeros(B, T, inner)
    for b in range(B):
        kb = k_all[b][idx[b]].view(T, S, heads, dim_head)  # gather per token
        vb = v_all[b][idx[b]].view(T, S, heads, dim_head)
        qb = q[b].view(T, heads, dim_head)
        sim = torch.einsum("thd,tshd->ths", qb, kb)
        pad = (idx[b] == V)  # [T,S]
        sim = sim + pad[:, None, :].float() * -1e4  # (:219-223)
        attn = sim.softmax(dim=-1)
        attn = attn * (~pad)[:, None, :].float()  # zeroed, not renormalised (:227-231)
        out[b] = torch.einsum("ths,tshd->thd", attn, vb).reshape(T, inner)
    return F.linear(out, sd[p + "to_out.weight"])


def gcp_block(x, vision, mask, sd, p="", heads=8, dim_head=64, return_gate=False):
    """GatedCrossAttentionBlock.forward (:347-374) under the shipped flags CONDITION_GATE, NONLINEAR_GATE, NO_CAT,
    FIX_ATTN_GATE=-1, no ADAPT layer (configs/pretrain/mq-glip-t.yaml:132-141)."""
    s = gcp_sparse_attention(x, vision, mask, sd, p + "attn.", heads, dim_head)
    g = _ln(s, sd, p + "attn_gate.norm")
    g = F.gelu(F.linear(g, sd[p + "attn_gate.linear1.weight"]))
    g = F.linear(g, sd[p + "attn_gate.linear2.weight"]).tanh()  # [B,T,1]
    x1 = s * g + x
    f = _ln(x1, sd, p + "ff.norm")
    f = F.linear(F.gelu(F.linear(f, sd[p + "ff.linear1.weight"])), sd[p + "ff.linear2.weight"])
    y = f * sd[p + "ff_gate"].tanh() + x1
    return (y, g) if return_gate else y


def dense_cross_attention(x, ctx, sd, p, heads, dim_head):
    """MaskedCrossAttention.forward with spase_forward=False, no mask (:196-248): x attends to ctx."""
    B, Tq, _ = x.shape
    xn = _ln(x, sd, p + "norm")
    cn = _ln(ctx, sd, p + "norm_kv")
    inner = heads * dim_head
    q = F.linear(xn, sd[p + "to_q.weight"]) * (dim_head ** -0.5)
    kv = F.linear(cn, sd[p + "to_kv.weight"])
    k, v = kv[..., :inner], kv[..., inner:]
    q = q.view(B, Tq, heads, dim_head).transpose(1, 2)
    k = k.view(B, -1, heads, dim_head).transpose(1, 2)
    v = v.view(B, -1, heads, dim_head).transpose(1, 2)
    attn = (q @ k.transpose(-1, -2)).softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B, Tq, inner)
    return F.linear(out, sd[p + "to_out.weight"])


def preselect(vision, image, sd, p="", heads=8, dim_head=32, num_layers=2, scale=1.0):
    """PreSelectModule.forward / PreSelectBlock.forward (:398-448): vision [B,V,256], image [B,I,256] -> [B,V,768]."""
    v = vision * scale
    img = image * scale
    for i in range(num_layers):
        lp = f"{p}layers.{i}."
        ca = dense_cross_attention(v, img, sd, lp + "image_condition.", heads, dim_head)
        res = F.linear(v, sd[lp + "res_mapping.weight"]) if (lp + "res_mapping.weight") in sd else v
        v = ca + res
        f = _ln(v, sd, lp + "ff.norm")
        v = F.linear(F.gelu(F.linear(f, sd[lp + "ff.linear1.weight"])), sd[lp + "ff.linear2.weight"]) + v
    return v


# ------------------------------------------------------------------------------------------------------------------
# BERT layers
# ------------------------------------------------------------------------------------------------------------------
def bert_layer(h, ext_mask, sd, p, heads=12, eps=1e-12, clamp=0.0):
    """Post-LN BERT layer: self-attention softmax(QK^T/sqrt(d) + ext_mask) -> dense + LN(. + h) -> GELU FFN + LN.
    HF-4.x BertLayer as called positionally by QVBertEncoder.forward (modeling_bert_new.py:600-608); third-party
    (transformers, un-pinned, requirements.txt:12) so restated from the in-repo copy
    maskrcnn_benchmark/modeling/rpn/modeling_bert.py:39-270 — whose +-5e4 clamps are applied when clamp>0."""
    B, T, D = h.shape
    d = D // heads

    def cl(t):
        return t.clamp(min=-clamp, max=clamp) if clamp > 0 else t

    q = _lin(h, sd, p + "attention.self.query").view(B, T, heads, d).transpose(1, 2)
    k = _lin(h, sd, p + "attention.self.key").view(B, T, heads, d).transpose(1, 2)
    v = _lin(h, sd, p + "attention.self.value").view(B, T, heads, d).transpose(1, 2)
    scores = cl((q @ k.transpose(-1, -2)) / math.sqrt(d))
    if ext_mask is not None:
        scores = scores + ext_mask
    ctx = (scores.softmax(-1) @ v).transpose(1, 2).reshape(B, T, D)
    # BertSelfOutput carries no clamp (rpn/modeling_bert.py:177-190); scores, BertIntermediate and BertOutput do
    a = _lin(ctx, sd, p + "attention.output.dense")
    a = _ln(a + h, sd, p + "attention.output.LayerNorm", eps)
    i = cl(F.gelu(cl(_lin(a, sd, p + "intermediate.dense"))))
    o = cl(_lin(i, sd, p + "output.dense"))
    return cl(_ln(o + a, sd, p + "output.LayerNorm", eps))


def extended_mask(attention_mask):
    """get_extended_attention_mask, HF-4.x classic form: (1 - mask) * -10000 broadcast to [B,1,1,T]."""
    return (1.0 - attention_mask[:, None, None, :].float()) * -10000.0


def bert_embeddings(input_ids, sd, p="embeddings.", eps=1e-12, position_ids=None):
    """QVBertEmbeddings.forward at eval (modeling_bert_new.py:457-519): word + token_type(0) + absolute position, LN.
    ``position_ids`` [B,T] (GroundingDINO restarts them per category, bertwarper.py:271-320) or None = arange(T)."""
    T = input_ids.shape[1]
    e = sd[p + "word_embeddings.weight"][input_ids]
    e = e + sd[p + "token_type_embeddings.weight"][0][None, None, :]
    e = e + (sd[p + "position_embeddings.weight"][:T][None] if position_ids is None else sd[p + "position_embeddings.weight"][position_ids])
    return _ln(e, sd, p + "LayerNorm", eps)


def qvbert_model(input_ids, attention_mask, vision, images, vision_mask, sd, num_layers=12, start_qv=6, heads=12):
    """QVBertModel.forward + QVBertEncoder.forward (modeling_bert_new.py:545-639,690-848) + BertEncoder.forward
    (bert_model_new.py:39-104, N_LAYERS=1).  GCP block i-6 is applied BEFORE BERT layer i for i >= 6."""
    h = bert_embeddings(input_ids, sd)
    ext = extended_mask(attention_mask)
    vq = None
    if vision is not None and images is not None and len(vision) > 0:
        vq = preselect(vision, images, sd, "pre_select.")
    for i in range(num_layers):
        if i >= start_qv and vq is not None:
            h = gcp_block(h, vq, vision_mask, sd, f"encoder.qv_layer.{i - start_qv}.")
        h = bert_layer(h, ext, sd, f"encoder.layer.{i}.", heads)
    m = attention_mask.float()
    embedded = h * m[..., None]
    aggregate = embedded.sum(1) / m.sum(-1, keepdim=True)
    return {"hidden": h, "embedded": embedded, "aggregate": aggregate, "masks": attention_mask, "vision": vq}


# ------------------------------------------------------------------------------------------------------------------
# VL deep fusion — maskrcnn_benchmark/utils/fuse_helper.py, maskrcnn_benchmark/modeling/rpn/vldyhead.py
# ------------------------------------------------------------------------------------------------------------------
def bi_attention(v, l, mask_l, sd, p="", heads=8, embed=2048):
    """BiAttentionBlockForCheckpoint.single_attention_call (:419-426) + BiMultiHeadAttention.forward (:218-303), eval.
    v [B,N,256] (levels concatenated), l [B,T,768], mask_l [B,T] (1 keep / 0 pad).  Returns (v', l')."""
    B, N, _ = v.shape
    T = l.shape[1]
    d = embed // heads
    vn = _ln(v, sd, p + "layer_norm_v")
    ln = _ln(l, sd, p + "layer_norm_l")
    q = _lin(vn, sd, p + "attn.v_proj") * (d ** -0.5)
    k = _lin(ln, sd, p + "attn.l_proj")
    vv = _lin(vn, sd, p + "attn.values_v_proj")
    vl = _lin(ln, sd, p + "attn.values_l_proj")

    def sh(t, n):
        return t.view(B, n, heads, d).transpose(1, 2)  # [B,h,n,d]

    q, k, vv, vl = sh(q, N), sh(k, T), sh(vv, N), sh(vl, T)
    # int64 mask: kept tokens add +1, padded tokens -9e15 (:270-283)
    m = torch.where(mask_l == 0, torch.full_like(mask_l, -9e15, dtype=torch.float32), mask_l.float())
    ov = v.new_empty(B, heads, N, d)
    ol = v.new_empty(B, heads, T, d)
    for b in range(B):  # one (image, head) score matrix at a time: same arithmetic, cache-sized working set
        for h in range(heads):
            A = (q[b, h] @ k[b, h].t()).clamp(min=-50000, max=50000)  # [N,T]
            At = A.t().contiguous()
            Al = (At - At.max(dim=-1, keepdim=True)[0]).clamp(min=-50000, max=50000).softmax(dim=-1)  # over N, no mask
            ol[b, h] = Al @ vv[b, h]
            ov[b, h] = (A + m[b][None, :]).softmax(dim=-1) @ vl[b, h]
    ov = ov.transpose(1, 2).reshape(B, N, embed)
    ol = ol.transpose(1, 2).reshape(B, T, embed)
    dv = _lin(ov, sd, p + "attn.out_v_proj")
    dl = _lin(ol, sd, p + "attn.out_l_proj")
    # residual on the NORMALISED inputs (:420-425)
    return vn + sd[p + "gamma_v"] * dv, ln + sd[p + "gamma_l"] * dl


def gdino_bi_attention(v, l, sd, p="", heads=4, embed=1024, mask_v=None, mask_l=None):
    """GroundingDINO's BiAttentionBlock.forward (groundingdino_new/models/GroundingDINO/fuse_modules.py:286-296) +
    BiMultiHeadAttention.forward (:147-254), eval: STABLE_SOFTMAX_2D (the global maximum of the whole score tensor is
    subtracted before the clamps, :177-178), boolean masks (True = padding) filled with -inf.
    v [B,N,v_dim], l [B,T,l_dim], mask_v [B,N] / mask_l [B,T] bool or None.  Returns (v', l')."""
    B, N, _ = v.shape
    T = l.shape[1]
    d = embed // heads
    vn = _ln(v, sd, p + "layer_norm_v")
    ln = _ln(l, sd, p + "layer_norm_l")
    q = _lin(vn, sd, p + "attn.v_proj") * (d ** -0.5)
    k = _lin(ln, sd, p + "attn.l_proj")
    vv = _lin(vn, sd, p + "attn.values_v_proj")
    vl = _lin(ln, sd, p + "attn.values_l_proj")

    def sh(t, n):
        return t.view(B, n, heads, d).transpose(1, 2)  # [B,h,n,d]

    q, k, vv, vl = sh(q, N), sh(k, T), sh(vv, N), sh(vl, T)
    A = q @ k.transpose(-1, -2)                                   # [B,h,N,T]
    A = (A - A.max()).clamp(min=-50000, max=50000)                # :177-187
    At = A.transpose(-1, -2)
    Al = (At - At.max(dim=-1, keepdim=True)[0]).clamp(min=-50000, max=50000)
    if mask_v is not None:
        Al = Al.masked_fill(mask_v[:, None, None, :], float("-inf"))
    Al = Al.softmax(dim=-1)
    if mask_l is not None:
        A = A.masked_fill(mask_l[:, None, None, :], float("-inf"))
    Av = A.softmax(dim=-1)
    ov = (Av @ vl).transpose(1, 2).reshape(B, N, embed)
    ol = (Al @ vv).transpose(1, 2).reshape(B, T, embed)
    dv = _lin(ov, sd, p + "attn.out_v_proj")
    dl = _lin(ol, sd, p + "attn.out_l_proj")
    return vn + sd[p + "gamma_v"] * dv, ln + sd[p + "gamma_l"] * dl


def gdino_two_stage_select(class_logits, coord_unsel, output_proposals, output_memory, k=900):
    """Two-stage query selection of the GroundingDINO transformer (transformer.py:297-318): the k proposals with the largest
    max-over-tokens class logit; gathers of their boxes (unsigmoid), anchor proposals (sigmoid) and memory rows.
    class_logits [B,Q,T] (may hold -inf on padded tokens), coord_unsel / output_proposals [B,Q,4], output_memory [B,Q,C]."""
    topk_logits = class_logits.max(-1)[0]
    idx = torch.topk(topk_logits, k, dim=1)[1]
    g4 = idx.unsqueeze(-1).repeat(1, 1, 4)
    refpoint = torch.gather(coord_unsel, 1, g4)
    init_box = torch.gather(output_proposals, 1, g4).sigmoid()
    tgt = torch.gather(output_memory, 1, idx.unsqueeze(-1).repeat(1, 1, output_memory.shape[-1]))
    return {"topk_logits": topk_logits, "idx": idx, "refpoint_embed": refpoint, "init_box_proposal": init_box, "tgt": tgt}


def dot_product_head(feat, hidden, sd, p=""):
    """VLDyHead.forward dot-product section (vldyhead.py:806-818,871-888): feat [B,N,256] (tower output, permuted
    and flattened), hidden [B,T,768] -> logits [B,N,T]."""
    e = F.normalize(hidden, p=2, dim=-1)
    tok = _lin(e / 2.0, sd, p + "dot_product_projection_text")  # [B,T,256]
    bias = e @ sd[p + "bias_lang"] + sd[p + "bias0"]  # [B,T]
    logit = (feat @ tok.transpose(-1, -2)) / sd[p + "log_scale"].exp() + bias[:, None, :]
    return logit.clamp(min=-50000, max=50000)


def contrastive_embed(x, encoded_text, text_token_mask, max_text_len=256):
    """GroundingDINO ContrastiveEmbed.forward (groundingdino_new/models/GroundingDINO/utils.py:242-268):
    x [B,Q,D] @ encoded_text [B,T,D]^T, -inf where the token is padding, right-padded with -inf to max_text_len."""
    res = x @ encoded_text.transpose(-1, -2)
    res = res.masked_fill(~text_token_mask[:, None, :].bool(), float("-inf"))
    out = torch.full((*res.shape[:-1], max_text_len), float("-inf"))
    out[..., : res.shape[-1]] = res
    return out


# ------------------------------------------------------------------------------------------------------------------
# Post-processing — maskrcnn_benchmark/modeling/rpn/inference.py, anchor_generator.py, csrc/cuda/ml_nms.cu
# ------------------------------------------------------------------------------------------------------------------
def anchors_level(grid_h, grid_w, stride, size):
    """AnchorGenerator.grid_anchors + generate_anchors for one square anchor per cell (anchor_generator.py:111-137,
    355-425): base window [0,0,s-1,s-1] scaled to `size` about its centre, shifted by (x*s, y*s)."""
    ctr = 0.5 * (stride - 1)
    ws = float(round(math.sqrt(stride * stride / 1.0)))  # _ratio_enum with ratio 1
    scale = size / stride
    w = ws * scale
    base = torch.tensor([ctr - 0.5 * (w - 1), ctr - 0.5 * (w - 1), ctr + 0.5 * (w - 1), ctr + 0.5 * (w - 1)])
    sx = torch.arange(0, grid_w * stride, stride, dtype=torch.float32)
    sy = torch.arange(0, grid_h * stride, stride, dtype=torch.float32)
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), dim=1)
    return shifts + base[None]


def box_decode(reg, anchors):
    """BoxCoder.decode (vldyhead.py:78-108), weights (10,10,5,5), TO_REMOVE=1, clamp log(1000/16)."""
    w = anchors[:, 2] - anchors[:, 0] + 1
    h = anchors[:, 3] - anchors[:, 1] + 1
    cx = (anchors[:, 2] + anchors[:, 0]) / 2
    cy = (anchors[:, 3] + anchors[:, 1]) / 2
    dx, dy = reg[:, 0] / 10.0, reg[:, 1] / 10.0
    dw = (reg[:, 2] / 5.0).clamp(max=math.log(1000.0 / 16))
    dh = (reg[:, 3] / 5.0).clamp(max=math.log(1000.0 / 16))
    pcx, pcy = dx * w + cx, dy * h + cy
    pw, ph = torch.exp(dw) * w, torch.exp(dh) * h
    return torch.stack((pcx - 0.5 * (pw - 1), pcy - 0.5 * (ph - 1), pcx + 0.5 * (pw - 1), pcy + 0.5 * (ph - 1)), dim=1)


def class_scores(logits, positive_map, num_classes):
    """sigmoid + convert_grounding_to_od_logits, MEAN aggregation (inference.py:651-684,772-790).
    logits [B,HW,T]; positive_map {label(1-based): [token idx]} -> scores [B,HW,C]."""
    p = logits.sigmoid()
    scores = torch.zeros(p.shape[0], p.shape[1], num_classes)
    for label, toks in positive_map.items():
        scores[:, :, label - 1] = p[:, :, torch.as_tensor(toks, dtype=torch.long)].mean(-1)
    return scores


def atss_level_candidates(logits, reg, ctr, anchors, positive_map, num_classes, img_w, img_h, pre_nms_thresh=0.05,
                          pre_nms_top_n=1000):
    """ATSSPostProcessor.forward_for_single_feature_map for ONE image (inference.py:620-712).
    logits [HW,T], reg [HW,4], ctr [HW], anchors [HW,4].  Returns dict(boxes, scores, labels, loc, cls) with the
    top-k candidates in canonical order (score desc, then (loc, cls) asc) — the reference's topk(sorted=False) order is
    implementation-defined."""
    sc = class_scores(logits[None], positive_map, num_classes)[0]  # [HW,C]
    cand = sc > pre_nms_thresh
    k = min(int(cand.sum()), pre_nms_top_n)
    s = sc * ctr.sigmoid()[:, None]
    loc, cls = torch.nonzero(cand, as_tuple=True)
    vals = s[cand]
    # exact top-k with deterministic tie-break (value desc, then location asc, then class asc): torch.nonzero already
    # enumerates (loc, cls) in ascending order, so one STABLE descending sort on the values is the lexicographic order
    order = torch.sort(vals, descending=True, stable=True)[1][:k]
    loc, cls, vals = loc[order], cls[order], vals[order]
    boxes = box_decode(reg[loc], anchors[loc])
    boxes[:, 0].clamp_(min=0, max=img_w - 1)
    boxes[:, 1].clamp_(min=0, max=img_h - 1)
    boxes[:, 2].clamp_(min=0, max=img_w - 1)
    boxes[:, 3].clamp_(min=0, max=img_h - 1)
    ws, hs = boxes[:, 2] - boxes[:, 0] + 1, boxes[:, 3] - boxes[:, 1] + 1  # remove_small_boxes(min_size=0)
    keep = (ws >= 0) & (hs >= 0)
    return {"boxes": boxes[keep], "scores": vals.sqrt()[keep], "labels": (cls + 1)[keep], "loc": loc[keep], "cls": cls[keep]}


def ml_nms(boxes, scores, labels, thresh, order=None):
    """maskrcnn_benchmark/csrc/cuda/ml_nms.cu: sort by score desc (:81-83), box i suppresses j>i iff same label and
    IoU(+1 widths) > thresh (:15-26,67), greedy in sorted order (:129-140), kept ORIGINAL indices ascending (:145-149).
    Plain fp32 arithmetic in numpy (no FMA contraction)."""
    import numpy as np

    n = boxes.shape[0]
    if n == 0:
        return torch.empty(0, dtype=torch.long)
    b = boxes.detach().cpu().numpy().astype(np.float32)
    s = scores.detach().cpu().numpy().astype(np.float32)
    lab = labels.detach().cpu().numpy().astype(np.float32)
    if order is None:
        order = np.argsort(-s.astype(np.float64), kind="stable")  # ties: lower index first
    else:
        order = order.cpu().numpy()
    b, lab = b[order], lab[order]
    one = np.float32(1.0)
    area = (b[:, 2] - b[:, 0] + one) * (b[:, 3] - b[:, 1] + one)
    removed = np.zeros(n, dtype=bool)
    keep = []
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        j = np.arange(i + 1, n)
        if j.size == 0:
            break
        left = np.maximum(b[i, 0], b[j, 0]); right = np.minimum(b[i, 2], b[j, 2])
        top = np.maximum(b[i, 1], b[j, 1]); bottom = np.minimum(b[i, 3], b[j, 3])
        w = np.maximum(right - left + one, np.float32(0)); h = np.maximum(bottom - top + one, np.float32(0))
        inter = (w * h).astype(np.float32)
        iou = inter / ((area[i] + area[j]).astype(np.float32) - inter)
        removed[j[(iou > np.float32(thresh)) & (lab[j] == lab[i])]] = True
    kept = np.sort(order[np.asarray(keep, dtype=np.int64)])
    return torch.from_numpy(kept.astype(np.int64))


def select_over_all_levels(boxes, scores, labels, nms_thresh=0.6, max_det=100):
    """ATSSPostProcessor.select_over_all_levels (inference.py:748-769): ml_nms then the kthvalue '>=' top-k cut."""
    keep = ml_nms(boxes, scores, labels.float(), nms_thresh)
    if keep.numel() > max_det > 0:
        ks = scores[keep]
        thr = torch.kthvalue(ks.float(), keep.numel() - max_det + 1)[0]
        keep = keep[ks >= thr]
    return keep


# ------------------------------------------------------------------------------------------------------------------
# DyConv — maskrcnn_benchmark/modeling/rpn/vldyhead.py:155-247, csrc/cuda/deform_conv_kernel_cuda.cu, layers/dyrelu.py
# ------------------------------------------------------------------------------------------------------------------
def dcn_v2(x, off_flat, mask_flat, weight, bias, stride, fast=False):
    """modulated_deform_conv_cuda_forward (deform_conv_cuda.cu:496-575) with the im2col of
    deform_conv_kernel_cuda.cu:578-641, 3x3, pad 1, dilation 1, one deformable group, per image.
    x [B,C,H,W]; off_flat [B, >=18*Ho*Wo] / mask_flat [B, >=9*Ho*Wo] are the NCHW-FLAT offset / (sigmoid-ed) mask buffers
    which the kernel indexes with the OUTPUT strides Ho*Wo (:605-618) — for DyConv[0] these buffers were produced at a
    finer level, i.e. they are re-interpreted."""
    B, C, H, W = x.shape
    Ho = (H + 2 - 3) // stride + 1
    Wo = (W + 2 - 3) // stride + 1
    P = Ho * Wo
    if fast and off_flat.shape[1] == 18 * P and mask_flat.shape[1] == 9 * P:
        # no re-interpretation involved: torchvision's deform_conv2d is the same algorithm (bit-identical to the
        # gather formulation below on the cases tests/test_oracle_golden.py checks) (kept as a cross-check; it is slower than the gather formulation on CPU)
        import torchvision
        return torchvision.ops.deform_conv2d(x, off_flat.view(B, 18, Ho, Wo), weight, bias, stride=stride, padding=1,
                                             mask=mask_flat.view(B, 9, Ho, Wo))
    ho = torch.arange(Ho).view(-1, 1).expand(Ho, Wo).reshape(-1)
    wo = torch.arange(Wo).view(1, -1).expand(Ho, Wo).reshape(-1)
    pix = torch.arange(P)
    xf = x.reshape(B, C, H * W)
    cols = x.new_zeros(B, C, 9, P)
    for t in range(9):
        i, j = t // 3, t % 3
        oh = off_flat[:, (2 * t) * P + pix]
        ow = off_flat[:, (2 * t + 1) * P + pix]
        m = mask_flat[:, t * P + pix]
        h_im = (ho * stride - 1 + i).float()[None] + oh
        w_im = (wo * stride - 1 + j).float()[None] + ow
        valid = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
        h_low, w_low = torch.floor(h_im), torch.floor(w_im)
        lh, lw = h_im - h_low, w_im - w_low
        hh, hw = 1 - lh, 1 - lw
        h_low, w_low = h_low.long(), w_low.long()
        h_high, w_high = h_low + 1, w_low + 1

        def corner(hc, wc, ok):
            idx = (hc.clamp(0, H - 1) * W + wc.clamp(0, W - 1))  # [B,P]
            v = torch.gather(xf, 2, idx[:, None, :].expand(B, C, P))
            return v * (ok & valid)[:, None, :].float()

        v1 = corner(h_low, w_low, (h_low >= 0) & (w_low >= 0))
        v2 = corner(h_low, w_high, (h_low >= 0) & (w_high <= W - 1))
        v3 = corner(h_high, w_low, (h_high <= H - 1) & (w_low >= 0))
        v4 = corner(h_high, w_high, (h_high <= H - 1) & (w_high <= W - 1))
        val = (hh * hw)[:, None] * v1 + (hh * lw)[:, None] * v2 + (lh * hw)[:, None] * v3 + (lh * lw)[:, None] * v4
        cols[:, :, t] = val * m[:, None, :]
    out = torch.einsum("bckp,ock->bop", cols, weight.reshape(weight.shape[0], C, 9)) + bias[None, :, None]
    return out.reshape(B, -1, Ho, Wo)


def _hsig(x):
    return F.relu6(x + 3.0) / 6.0


def dyrelu(x, sd, p):
    """DYReLU.forward (layers/dyrelu.py:80-104), K2 + bias, lambda_a = 2, init_a = [1, 0], init_b = [0, 0]."""
    B, C, _, _ = x.shape
    y = x.mean(dim=(2, 3))
    y = F.relu(F.linear(y, sd[p + "fc.0.weight"], sd[p + "fc.0.bias"]))
    y = _hsig(F.linear(y, sd[p + "fc.2.weight"], sd[p + "fc.2.bias"])).view(B, 4 * C, 1, 1)
    a1, b1, a2, b2 = torch.split(y, C, dim=1)
    a1 = (a1 - 0.5) * 2.0 + 1.0
    a2 = (a2 - 0.5) * 2.0 + 0.0
    b1 = b1 - 0.5
    b2 = b2 - 0.5
    return torch.max(x * a1 + b1, x * a2 + b2)


def dyconv(feats, sd, p=""):
    """DyConv.forward (vldyhead.py:205-247): feats = list of [B,256,h,w]; returns the next pyramid."""
    L = len(feats)
    B = feats[0].shape[0]
    nxt = []
    for l, f in enumerate(feats):
        om = F.conv2d(f, sd[p + "offset.weight"], sd[p + "offset.bias"], padding=1)
        off_flat = om[:, :18].reshape(B, -1)  # offset[b] is a contiguous [18,H,W] block of the 27-channel buffer
        mask_flat = om[:, 18:].sigmoid().reshape(B, -1)  # .sigmoid() materialises a contiguous [9,H,W] tensor

        def conv(k, x, stride):
            y = dcn_v2(x, off_flat, mask_flat, sd[f"{p}DyConv.{k}.conv.weight"], sd[f"{p}DyConv.{k}.conv.bias"], stride)
            return F.group_norm(y, 16, sd[f"{p}DyConv.{k}.bn.weight"], sd[f"{p}DyConv.{k}.bn.bias"], 1e-5)

        temp = [conv(1, f, 1)]
        if l > 0:
            temp.append(conv(2, feats[l - 1], 2))
        if l < L - 1:
            temp.append(F.interpolate(conv(0, feats[l + 1], 1), size=f.shape[2:], mode="bilinear", align_corners=True))
        attn = []
        for t in temp:
            a = F.relu(F.conv2d(t.mean(dim=(2, 3), keepdim=True), sd[p + "AttnConv.1.weight"], sd[p + "AttnConv.1.bias"]))
            attn.append(_hsig(a))
        mean_fea = torch.stack([t * a for t, a in zip(temp, attn)]).mean(dim=0)
        nxt.append(mean_fea)
    return [dyrelu(t, sd, p + "relu.") for t in nxt]


def flatten_levels(feats):
    """permute_and_flatten + cat (fuse_helper.py:398-404): [B,C,h,w] x L -> [B, N, C]."""
    return torch.cat([f.flatten(2).transpose(1, 2) for f in feats], dim=1)


def split_levels(v, sizes):
    out, s = [], 0
    B, _, C = v.shape
    for h, w in sizes:
        out.append(v[:, s:s + h * w].transpose(1, 2).reshape(B, C, h, w))
        s += h * w
    return out


def vl_dyhead(feats, hidden, masks, sd, num_convs=6):
    """VLDyHead.forward (vldyhead.py:769-900) for the MQ-GLIP configuration: per layer VLFuse(MHA-B) ->
    BertEncoderLayer -> DyConv; then the dot-product token head and the 1x1 bbox / centerness heads (+ Scale)."""
    sizes = [f.shape[2:] for f in feats]
    for i in range(num_convs):
        v, hidden = bi_attention(flatten_levels(feats), hidden, masks, sd, f"dyhead_tower.{3 * i}.b_attn.")
        feats = split_levels(v, sizes)
        hidden = bert_layer(hidden, extended_mask(masks), sd, f"dyhead_tower.{3 * i + 1}.", clamp=50000.0)
        feats = dyconv(feats, sd, f"dyhead_tower.{3 * i + 2}.")
    v = flatten_levels(feats)
    logits = dot_product_head(v, hidden, sd)
    bbox, ctr = [], []
    for l, f in enumerate(feats):
        bbox.append(F.conv2d(f, sd["bbox_pred.weight"], sd["bbox_pred.bias"]) * sd[f"scales.{l}.scale"])
        ctr.append(F.conv2d(f, sd["centerness.weight"], sd["centerness.bias"]))
    return {"dot_product_logits": logits, "bbox_reg": bbox, "centerness": ctr, "visual": feats, "hidden": hidden}


# ------------------------------------------------------------------------------------------------------------------
# Swin backbone + FPN — maskrcnn_benchmark/modeling/backbone/swint.py, fpn.py
# ------------------------------------------------------------------------------------------------------------------
def _rel_pos_index(ws):
    """pair-wise relative position index of a ws x ws window (swint.py:88-99)."""
    idx = torch.arange(ws)
    dh = idx[:, None, None, None] - idx[None, None, :, None] + ws - 1  # [ih, iw, jh, jw] -> only ih, jh vary
    dw = idx[None, :, None, None] - idx[None, None, None, :] + ws - 1
    return (dh * (2 * ws - 1) + dw).expand(ws, ws, ws, ws).reshape(ws * ws, ws * ws)


def swin_block(x, H, W, sd, p, heads, ws, shift):
    """SwinTransformerBlock.forward (swint.py:186-242) + WindowAttention.forward (:111-142) + the shift mask of
    BasicLayer.forward (:354-373).  x [B, H*W, C]."""
    B, L, C = x.shape
    d = C // heads
    short = x
    xn = _ln(x, sd, p + "norm1").view(B, H, W, C)
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    xn = F.pad(xn, (0, 0, 0, Wp - W, 0, Hp - H))  # zero padding AFTER the norm (:200-205)
    if shift > 0:
        xn = torch.roll(xn, shifts=(-shift, -shift), dims=(1, 2))
        ids = torch.zeros(Hp, Wp)
        bounds_h = [(0, Hp - ws), (Hp - ws, Hp - shift), (Hp - shift, Hp)]
        bounds_w = [(0, Wp - ws), (Wp - ws, Wp - shift), (Wp - shift, Wp)]
        cnt = 0
        for h0, h1 in bounds_h:
            for w0, w1 in bounds_w:
                ids[h0:h1, w0:w1] = cnt
                cnt += 1
        mw = ids.view(Hp // ws, ws, Wp // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
        mask = (mw[:, None, :] != mw[:, :, None]).float() * -100.0  # [nW, N, N]
    else:
        mask = None
    win = xn.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)
    qkv = _lin(win, sd, p + "attn.qkv").view(-1, ws * ws, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * d ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    bias = sd[p + "attn.relative_position_bias_table"][_rel_pos_index(ws).reshape(-1)].view(ws * ws, ws * ws, heads)
    attn = attn + bias.permute(2, 0, 1)[None]
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B, nW, heads, ws * ws, ws * ws) + mask[None, :, None]).view(-1, heads, ws * ws, ws * ws)
    o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(-1, ws * ws, C)
    o = _lin(o, sd, p + "attn.proj")
    o = o.view(B, Hp // ws, Wp // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
    if shift > 0:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    x = short + o[:, :H, :W].reshape(B, L, C)
    m = _lin(F.gelu(_lin(_ln(x, sd, p + "norm2"), sd, p + "mlp.fc1")), sd, p + "mlp.fc2")
    return x + m


def swin_transformer(img, sd, depths=(2, 2, 6, 2), heads=(3, 6, 12, 24), embed=96, ws=7, p=""):
    """SwinTransformer.forward (swint.py:591-615): returns the 4 normalised stage outputs as [B, C, H, W]."""
    img = F.pad(img, (0, (-img.shape[3]) % 4, 0, (-img.shape[2]) % 4))  # PatchEmbed zero-pads to a multiple of 4 (:413-418)
    x = F.conv2d(img, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=4)
    B, _, H, W = x.shape
    x = _ln(x.flatten(2).transpose(1, 2), sd, p + "patch_embed.norm")
    outs = []
    for i, depth in enumerate(depths):
        C = embed * 2 ** i
        for j in range(depth):
            x = swin_block(x, H, W, sd, f"{p}layers.{i}.blocks.{j}.", heads[i], ws, 0 if j % 2 == 0 else ws // 2)
        y = _ln(x, sd, f"{p}norm{i}") if i > 0 else x  # norm0 is nn.Identity for *-RETINANET backbones (:547-548)
        outs.append(y.view(B, H, W, C).permute(0, 3, 1, 2))
        if i < len(depths) - 1:  # PatchMerging (:256-284)
            g = F.pad(x.view(B, H, W, C), (0, 0, 0, W % 2, 0, H % 2))
            g = torch.cat([g[:, 0::2, 0::2], g[:, 1::2, 0::2], g[:, 0::2, 1::2], g[:, 1::2, 1::2]], -1)
            H, W = (H + 1) // 2, (W + 1) // 2
            x = F.linear(_ln(g.view(B, H * W, 4 * C), sd, f"{p}layers.{i}.downsample.norm"),
                         sd[f"{p}layers.{i}.downsample.reduction.weight"])
    return outs


def fpn(stage_outs, sd, p=""):
    """FPN.forward (fpn.py:59-129) for SWINT-FPN-RETINANET: laterals on C3..C5, nearest top-down, 3x3 output convs,
    P6 = conv_s2(P5), P7 = conv_s2(relu(P6)).  stage_outs: the 4 Swin outputs (the first is unused)."""
    c3, c4, c5 = stage_outs[-3:]

    def conv(name, t, **kw):
        return F.conv2d(t, sd[p + name + ".weight"], sd[p + name + ".bias"], **kw)

    i5 = conv("fpn_inner4", c5)
    p5 = conv("fpn_layer4", i5, padding=1)
    i4 = conv("fpn_inner3", c4) + F.interpolate(i5, size=c4.shape[-2:], mode="nearest")
    p4 = conv("fpn_layer3", i4, padding=1)
    i3 = conv("fpn_inner2", c3) + F.interpolate(i4, size=c3.shape[-2:], mode="nearest")
    p3 = conv("fpn_layer2", i3, padding=1)
    p6 = conv("top_blocks.p6", p5, stride=2, padding=1)
    p7 = conv("top_blocks.p7", F.relu(p6), stride=2, padding=1)
    return [p3, p4, p5, p6, p7]


# ------------------------------------------------------------------------------------------------------------------
# GroundingDINO multi-scale deformable attention — groundingdino_new/models/GroundingDINO/ms_deform_attn.py:93-133,236-352
# ------------------------------------------------------------------------------------------------------------------
def ms_deform_attn(query, value, reference_points, spatial_shapes, sd, p="", heads=8, points=4, key_padding_mask=None,
                   query_pos=None):
    """MultiScaleDeformableAttention.forward with batch_first=True: query [B,Q,E], value [B,Nv,E], reference_points
    [B,Q,L,2|4] -> [B,Q,E]; the sampling core is the reference's own CPU path (grid_sample, :93-133)."""
    if query_pos is not None:
        query = query + query_pos
    B, Q, E = query.shape
    Nv = value.shape[1]
    L = len(spatial_shapes)
    d = E // heads
    v = _lin(value, sd, p + "value_proj")
    if key_padding_mask is not None:
        v = v.masked_fill(key_padding_mask[..., None], 0.0)
    v = v.view(B, Nv, heads, d)
    off = _lin(query, sd, p + "sampling_offsets").view(B, Q, heads, L, points, 2)
    aw = _lin(query, sd, p + "attention_weights").view(B, Q, heads, L * points).softmax(-1).view(B, Q, heads, L, points)
    ss = torch.tensor(spatial_shapes, dtype=torch.float32)
    if reference_points.shape[-1] == 2:
        norm = torch.stack([ss[:, 1], ss[:, 0]], -1)
        loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = reference_points[:, :, None, :, None, :2] + off / points * reference_points[:, :, None, :, None, 2:] * 0.5
    vals = v.split([h * w for h, w in spatial_shapes], dim=1)
    grids = 2 * loc - 1
    sampled = []
    for l, (h, w) in enumerate(spatial_shapes):
        vl = vals[l].flatten(2).transpose(1, 2).reshape(B * heads, d, h, w)
        g = grids[:, :, :, l].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(vl, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    aw = aw.transpose(1, 2).reshape(B * heads, 1, Q, L * points)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * aw).sum(-1).view(B, heads * d, Q).transpose(1, 2)
    return _lin(out, sd, p + "output_proj")


# ------------------------------------------------------------------------------------------------------------------
# Vision-query extraction — generalized_vl_rcnn_new.py:232-288, modeling/poolers.py:11-129, layers/roi_align.py:71-81
# ------------------------------------------------------------------------------------------------------------------
def expand_boxes(bbox, image_size, ratio=1.5):
    """expand_bbox (:32-49): grow about the centre, clip to [0, w-1] x [0, h-1], drop empty boxes.  Returns (boxes, keep)."""
    w, h = image_size
    bw, bh = bbox[:, 2] - bbox[:, 0], bbox[:, 3] - bbox[:, 1]
    dw, dh = (bw * ratio - bw) / 2, (bh * ratio - bh) / 2
    nb = bbox + torch.stack([-dw, -dh, dw, dh], dim=1)
    nb[:, 0].clamp_(min=0, max=w - 1); nb[:, 1].clamp_(min=0, max=h - 1)
    nb[:, 2].clamp_(min=0, max=w - 1); nb[:, 3].clamp_(min=0, max=h - 1)
    keep = (nb[:, 3] > nb[:, 1]) & (nb[:, 2] > nb[:, 0])
    return nb[keep], keep


def pool_query_features(pyr, boxes_per_image, scales=(0.125, 0.0625, 0.03125, 0.015625, 0.0078125), resolution=7,
                        sampling_ratio=0):
    """Pooler.forward with use_v2=True (poolers.py:99-129) + mean over the bins (:263): pyr = list of [B,C,h,w] fp32 maps,
    boxes_per_image = list of [K_i, 4] xyxy boxes (already expanded) -> ([sum K, C] features, [sum K] level indices)."""
    import math
    from torchvision.ops import roi_align
    rois = torch.cat([torch.cat([torch.full((b.shape[0], 1), float(i)), b.float()], dim=1) for i, b in enumerate(boxes_per_image)])
    k_min, k_max = -math.log2(scales[0]), -math.log2(scales[-1])
    area = (rois[:, 3] - rois[:, 1] + 1) * (rois[:, 4] - rois[:, 2] + 1)          # BoxList.area(), TO_REMOVE = 1
    lvls = torch.floor(4 + torch.log2(torch.sqrt(area) / 224 + 1e-6)).clamp(min=k_min, max=k_max).to(torch.int64) - int(k_min)
    out = torch.zeros((rois.shape[0], pyr[0].shape[1], resolution, resolution))
    for l, (f, sc) in enumerate(zip(pyr, scales)):
        idx = torch.nonzero(lvls == l).squeeze(1)
        if idx.numel():
            out[idx] = roi_align(f, rois[idx], (resolution, resolution), sc, sampling_ratio, aligned=True)
    return out.mean(dim=[-2, -1]), lvls


# ------------------------------------------------------------------------------------------------------------------
# Whole forward — maskrcnn_benchmark/modeling/detector/generalized_vl_rcnn_new.py:307-519 (eval), rpn/vldyhead.py:933-989
# ------------------------------------------------------------------------------------------------------------------
def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def detector(img, image_size, input_ids, attention_mask, positive_map, bank, sd, K=5, num_classes=80,
             strides=(8, 16, 32, 64, 128), anchor_sizes=(64, 128, 256, 512, 1024), pre_nms_thresh=0.05, pre_nms_top_n=1000,
             nms_thresh=0.6, max_det=100):
    """GeneralizedVLRCNN_New.forward at eval for a batch sharing one prompt.  img [B,3,H,W] (already normalised and
    padded), image_size (h, w) of the un-padded images, bank {label: [n, 1, 256]} with n == K (so the selector's
    sorted random choice is the identity).  Returns intermediates + per-image (boxes, scores, labels)."""
    B = img.shape[0]
    T = input_ids.shape[1]
    c = swin_transformer(img, _sub(sd, "backbone.body."))
    pyr = fpn(c, _sub(sd, "backbone.fpn."))
    labels = [k for k, v in positive_map.items() if len(v) != 0]
    vision = torch.cat([bank[l][:K].flatten(0, 1) for l in labels])[None].expand(B, -1, -1)
    vmask = torch.zeros(1, vision.shape[1], T)
    r = 0
    for l in labels:
        n = bank[l][:K].flatten(0, 1).shape[0]
        vmask[0, r:r + n, positive_map[l]] = 1.0
        r += n
    vmask = vmask.expand(B, -1, -1)
    pooled = torch.cat([F.avg_pool2d(f, 2).flatten(2) for f in pyr], dim=2).permute(0, 2, 1)
    ids = input_ids.expand(B, -1) if input_ids.shape[0] == 1 else input_ids
    am = attention_mask.expand(B, -1) if attention_mask.shape[0] == 1 else attention_mask
    lang = qvbert_model(ids, am, vision, pooled, vmask, _sub(sd, "language_backbone.body.model."))
    head = vl_dyhead(pyr, lang["hidden"], am, _sub(sd, "rpn.head."))
    ih, iw = image_size
    dets = []
    for b in range(B):
        boxes, scores, labs = [], [], []
        off = 0
        for l, f in enumerate(pyr):
            h, w = f.shape[2:]
            anchors = anchors_level(h, w, strides[l], anchor_sizes[l])
            reg = head["bbox_reg"][l][b].permute(1, 2, 0).reshape(-1, 4)
            ctr = head["centerness"][l][b].reshape(-1)
            cand = atss_level_candidates(head["dot_product_logits"][b, off:off + h * w], reg, ctr, anchors, positive_map,
                                         num_classes, iw, ih, pre_nms_thresh, pre_nms_top_n)
            boxes.append(cand["boxes"]); scores.append(cand["scores"]); labs.append(cand["labels"])
            off += h * w
        boxes, scores, labs = torch.cat(boxes), torch.cat(scores), torch.cat(labs)
        keep = select_over_all_levels(boxes, scores, labs, nms_thresh, max_det)
        dets.append((boxes[keep], scores[keep], labs[keep]))
    return {"pyramid": pyr, "hidden": lang["hidden"], "vision": lang["vision"], "logits": head["dot_product_logits"],
            "fused_hidden": head["hidden"], "detections": dets}


# ------------------------------------------------------------------------------------------------------------------
# GroundingDINO transformer and forward — groundingdino_new/models/GroundingDINO/{transformer,transformer_vanilla,utils,
# groundingdino,bertwarper}.py, backbone/{backbone,position_encoding}.py   (MQ-GroundingDINO-T, BASELINE config 4)
# ------------------------------------------------------------------------------------------------------------------
def sine_pos_embed(pos_tensor, num_pos_feats=128, temperature=10000, exchange_xy=True):
    """get_sine_pos_embed (utils.py:24-56): pos_tensor [B, n, k] -> [B, n, k * num_pos_feats]."""
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    res = []
    for x in pos_tensor.split([1] * pos_tensor.shape[-1], dim=-1):
        s = x * (2 * math.pi) / dim_t
        res.append(torch.stack((s[..., 0::2].sin(), s[..., 1::2].cos()), dim=3).flatten(2))
    if exchange_xy:
        res[0], res[1] = res[1], res[0]
    return torch.cat(res, dim=-1)


def sineembed_for_position(pos):
    """gen_sineembed_for_position (utils.py:203-232) for pos [..., 4] (x, y, w, h) -> [..., 512] ordered (y, x, w, h)."""
    dim_t = torch.arange(128, dtype=torch.float32)
    dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / 128)
    out = []
    for c in (1, 0, 2, 3):
        e = pos[..., c, None] * (2 * math.pi) / dim_t
        out.append(torch.stack((e[..., 0::2].sin(), e[..., 1::2].cos()), dim=-1).flatten(-2))
    return torch.cat(out, dim=-1)


def position_embedding_sine_hw(mask, num_pos_feats=128, temperature_h=20, temperature_w=20):
    """PositionEmbeddingSineHW.forward with normalize=True (backbone/position_encoding.py:78-128, build at :171-178):
    mask [B,H,W] bool (True = padding) -> [B, 2*num_pos_feats, H, W]."""
    not_mask = ~mask
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    eps, scale = 1e-6, 2 * math.pi
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    d = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_tx = temperature_w ** (2 * torch.div(d, 2, rounding_mode="floor") / num_pos_feats)
    dim_ty = temperature_h ** (2 * torch.div(d, 2, rounding_mode="floor") / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_tx
    pos_y = y_embed[:, :, :, None] / dim_ty
    pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def mha(q_in, k_in, v_in, sd, p, heads, attn_mask=None, key_padding_mask=None):
    """torch.nn.MultiheadAttention.forward at eval (third-party: torch; call sites transformer.py:843,851,
    transformer_vanilla.py:112), batch-first restatement: q_in [B,Lq,E], k_in / v_in [B,Lk,E]; attn_mask bool [B,Lq,Lk]
    (True = not allowed, shared by the heads), key_padding_mask bool [B,Lk] (True = ignored) -> [B,Lq,E]."""
    B, Lq, E = q_in.shape
    Lk = k_in.shape[1]
    d = E // heads
    w, b = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    q = F.linear(q_in, w[:E], b[:E]).view(B, Lq, heads, d).transpose(1, 2) * (d ** -0.5)
    k = F.linear(k_in, w[E:2 * E], b[E:2 * E]).view(B, Lk, heads, d).transpose(1, 2)
    v = F.linear(v_in, w[2 * E:], b[2 * E:]).view(B, Lk, heads, d).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if attn_mask is not None:
        s = s.masked_fill(attn_mask[:, None], float("-inf"))
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    o = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, Lq, E)
    return _lin(o, sd, p + "out_proj")


def gdino_text_enhancer_layer(src, pos, attn_mask, sd, p, heads=4):
    """TransformerEncoderLayer.forward (transformer_vanilla.py:96-123), post-norm, relu; the key padding mask is NOT used
    (:114 is commented out in the reference).  src / pos [B,T,256], attn_mask bool [B,T,T] (True = not allowed)."""
    q = src + pos
    src = _ln(src + mha(q, q, src, sd, p + "self_attn.", heads, attn_mask=attn_mask), sd, p + "norm1")
    ff = _lin(F.relu(_lin(src, sd, p + "linear1")), sd, p + "linear2")
    return _ln(src + ff, sd, p + "norm2")


def gdino_deformable_encoder_layer(src, pos, reference_points, spatial_shapes, key_padding_mask, sd, p, heads=8, points=4):
    """DeformableTransformerEncoderLayer.forward (transformer.py:729-760)."""
    src2 = ms_deform_attn(src + pos, src, reference_points, spatial_shapes, sd, p + "self_attn.", heads, points,
                          key_padding_mask=key_padding_mask)
    src = _ln(src + src2, sd, p + "norm1")
    ff = _lin(F.relu(_lin(src, sd, p + "linear1")), sd, p + "linear2")
    return _ln(src + ff, sd, p + "norm2")


def gdino_encoder_reference_points(spatial_shapes, valid_ratios):
    """TransformerEncoder.get_reference_points (transformer.py:473-489): [B, sum(hw), L, 2]."""
    refs = []
    for lvl, (H_, W_) in enumerate(spatial_shapes):
        ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H_ - 0.5, H_), torch.linspace(0.5, W_ - 0.5, W_), indexing="ij")
        ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
        ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
        refs.append(torch.stack((ref_x, ref_y), -1))
    return torch.cat(refs, 1)[:, :, None] * valid_ratios[:, None]


def gdino_valid_ratio(mask):
    """Transformer.get_valid_ratio (transformer.py:194-201): mask [B,H,W] bool -> [B, 2] (w, h)."""
    _, H, W = mask.shape
    return torch.stack([(~mask[:, 0, :]).sum(1).float() / W, (~mask[:, :, 0]).sum(1).float() / H], -1)


def gdino_encoder_output_proposals(memory, memory_padding_mask, spatial_shapes):
    """gen_encoder_output_proposals (utils.py:59-118), learnedwh=None."""
    N_ = memory.shape[0]
    proposals, cur = [], 0
    for lvl, (H_, W_) in enumerate(spatial_shapes):
        m = memory_padding_mask[:, cur:cur + H_ * W_].view(N_, H_, W_, 1)
        valid_H = torch.sum(~m[:, :, 0, 0], 1)
        valid_W = torch.sum(~m[:, 0, :, 0], 1)
        gy, gx = torch.meshgrid(torch.linspace(0, H_ - 1, H_), torch.linspace(0, W_ - 1, W_), indexing="ij")
        grid = torch.cat([gx.unsqueeze(-1), gy.unsqueeze(-1)], -1)
        scale = torch.cat([valid_W.unsqueeze(-1), valid_H.unsqueeze(-1)], 1).view(N_, 1, 1, 2)
        grid = (grid.unsqueeze(0).expand(N_, -1, -1, -1) + 0.5) / scale
        wh = torch.ones_like(grid) * 0.05 * (2.0 ** lvl)
        proposals.append(torch.cat((grid, wh), -1).view(N_, -1, 4))
        cur += H_ * W_
    prop = torch.cat(proposals, 1)
    valid = ((prop > 0.01) & (prop < 0.99)).all(-1, keepdim=True)
    prop = torch.log(prop / (1 - prop))
    prop = prop.masked_fill(memory_padding_mask.unsqueeze(-1), float("inf")).masked_fill(~valid, float("inf"))
    mem = memory.masked_fill(memory_padding_mask.unsqueeze(-1), 0.0).masked_fill(~valid, 0.0)
    return mem, prop


def _mlp(x, sd, p, n):
    """MLP (utils.py:171-185): Linear + relu ... Linear."""
    for i in range(n):
        x = _lin(x, sd, f"{p}layers.{i}")
        if i < n - 1:
            x = F.relu(x)
    return x


def inverse_sigmoid(x, eps=1e-3):
    """groundingdino_new/util/misc.py:721-725."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def gdino_decoder_layer(tgt, query_pos, ref_input, memory, memory_mask, spatial_shapes, memory_text, text_pad_mask, sd, p,
                        heads=8, points=4):
    """DeformableTransformerDecoderLayer.forward (transformer.py:823-878), batch-first: self-attention over the queries,
    text cross-attention, deformable cross-attention into the image memory, FFN — each followed by residual + LayerNorm."""
    q = tgt + query_pos
    tgt = _ln(tgt + mha(q, q, tgt, sd, p + "self_attn.", heads), sd, p + "norm2")
    tgt = _ln(tgt + mha(tgt + query_pos, memory_text, memory_text, sd, p + "ca_text.", heads,
                        key_padding_mask=text_pad_mask), sd, p + "catext_norm")
    tgt2 = ms_deform_attn(tgt + query_pos, memory, ref_input, spatial_shapes, sd, p + "cross_attn.", heads, points,
                          key_padding_mask=memory_mask)
    tgt = _ln(tgt + tgt2, sd, p + "norm1")
    ff = _lin(F.relu(_lin(tgt, sd, p + "linear1")), sd, p + "linear2")
    return _ln(tgt + ff, sd, p + "norm3")


def gdino_transformer(srcs, masks, poss, encoded_text, text_token_mask, position_ids, text_self_attention_masks, sd,
                      num_queries=900, enc_layers=6, dec_layers=6, heads=8, points=4, return_all=False):
    """Transformer.forward (transformer.py:206-403) with two_stage_type="standard", embed_init_tgt=True, text enhancer,
    fusion layers and text cross-attention (the shipped MQ-GroundingDINO-T settings, config/defaults.py:944-987), eval.
    srcs / poss: lists of [B,256,h,w]; masks: list of bool [B,h,w] (True = padding); encoded_text [B,T,256];
    text_token_mask bool [B,T] (True = token in use); position_ids int64 [B,T]; text_self_attention_masks bool [B,T,T]
    (True = may attend).  ``sd`` holds the Transformer's state_dict plus ``enc_out_bbox_embed.*`` and
    ``decoder.bbox_embed.{i}.*`` as the reference registers them (groundingdino.py:255-275).
    Returns hs (list of [B,nq,256], normed) and references (list of dec_layers+1 sigmoid boxes [B,nq,4])."""
    spatial_shapes = [tuple(s.shape[-2:]) for s in srcs]
    B = srcs[0].shape[0]
    src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    mask = torch.cat([m.flatten(1) for m in masks], 1)
    pos = torch.cat([p_.flatten(2).transpose(1, 2) + sd["level_embed"][l].view(1, 1, -1) for l, p_ in enumerate(poss)], 1)
    valid_ratios = torch.stack([gdino_valid_ratio(m) for m in masks], 1)
    # ---- encoder (transformer.py:491-590)
    ref_enc = gdino_encoder_reference_points(spatial_shapes, valid_ratios)
    pos_text = sine_pos_embed(position_ids[..., None].float(), num_pos_feats=256, exchange_xy=False)
    text_pad = ~text_token_mask
    out, text = src, encoded_text
    for i in range(enc_layers):
        out, text = gdino_bi_attention(out, text, sd, f"encoder.fusion_layers.{i}.", heads=heads // 2, embed=1024,
                                       mask_v=mask, mask_l=text_pad)
        text = gdino_text_enhancer_layer(text, pos_text, ~text_self_attention_masks, sd, f"encoder.text_layers.{i}.",
                                         heads=heads // 2)
        out = gdino_deformable_encoder_layer(out, pos, ref_enc, spatial_shapes, mask, sd, f"encoder.layers.{i}.", heads, points)
    memory, memory_text = out, text
    # ---- two-stage query selection (:272-318)
    output_memory, output_proposals = gdino_encoder_output_proposals(memory, mask, spatial_shapes)
    output_memory = _ln(_lin(output_memory, sd, "enc_output"), sd, "enc_output_norm")
    enc_class = contrastive_embed(output_memory, memory_text, text_token_mask)
    enc_coord = _mlp(output_memory, sd, "enc_out_bbox_embed.", 3) + output_proposals
    sel = gdino_two_stage_select(enc_class, enc_coord, output_proposals, output_memory, num_queries)
    refpoint = sel["refpoint_embed"]
    tgt = sd["tgt_embed.weight"][None].expand(B, -1, -1)
    # ---- decoder (:636-716)
    reference_points = refpoint.sigmoid()
    refs, hs = [reference_points], []
    vr4 = torch.cat([valid_ratios, valid_ratios], -1)
    output = tgt
    for i in range(dec_layers):
        ref_input = reference_points[:, :, None] * vr4[:, None]
        sine = sineembed_for_position(ref_input[:, :, 0, :])
        query_pos = _mlp(sine, sd, "decoder.ref_point_head.", 2)
        output = gdino_decoder_layer(output, query_pos, ref_input, memory, mask, spatial_shapes, memory_text, text_pad, sd,
                                     f"decoder.layers.{i}.", heads, points)
        delta = _mlp(output, sd, f"decoder.bbox_embed.{i}.", 3)
        reference_points = (delta + inverse_sigmoid(reference_points)).sigmoid()
        refs.append(reference_points)
        hs.append(_ln(output, sd, "decoder.norm"))
    res = {"hs": hs, "references": refs, "memory": memory, "memory_text": memory_text}
    if return_all:
        res.update(enc_class=enc_class, topk=sel["idx"], output_memory=output_memory, output_proposals=output_proposals)
    return res


def gdino_text_masks(input_ids, special_tokens=(101, 102, 1012, 1029)):
    """generate_masks_with_special_tokens_and_transfer_map (bertwarper.py:271-320): per-category block-diagonal self-attention
    mask bool [B,T,T] (True = may attend) and position ids that restart after every delimiter ([CLS], [SEP], '.', '?')."""
    bs, T = input_ids.shape
    special = torch.zeros((bs, T), dtype=torch.bool)
    for s in special_tokens:
        special |= input_ids == s
    mask = torch.eye(T, dtype=torch.bool)[None].repeat(bs, 1, 1)
    pid = torch.zeros((bs, T), dtype=torch.long)
    prev = 0
    for row, col in torch.nonzero(special).tolist():
        if col == 0 or col == T - 1:
            mask[row, col, col] = True
            pid[row, col] = 0
        else:
            mask[row, prev + 1:col + 1, prev + 1:col + 1] = True
            pid[row, prev + 1:col + 1] = torch.arange(0, col - prev)
        prev = col
    return mask, pid


def gdino_detections(pred_logits, pred_boxes, positive_map, num_classes, image_sizes, box_threshold=0.05):
    """convert_groundingdino_to_glip_output (groundingdino.py:291-335) on RAW class logits: sigmoid, MEAN score aggregation
    (rpn/inference.py:772-790), best class above the box threshold, cxcywh -> xyxy pixels, clip_to_image (TO_REMOVE = 1),
    remove_small_boxes(min_size=0).  Returns per image (boxes [k,4], scores [k], labels [k])."""
    prob = pred_logits.sigmoid()
    B, N, _ = prob.shape
    scores = torch.zeros(B, N, num_classes)
    for label, toks in positive_map.items():
        scores[:, :, label - 1] = prob[:, :, torch.as_tensor(list(toks), dtype=torch.long)].mean(-1)
    out = []
    for b, (H, W) in enumerate(image_sizes):
        cand = scores[b].max(-1)[0] > box_threshold
        s, idx = scores[b][cand].max(-1)
        box = pred_boxes[b][cand] * torch.tensor([W, H, W, H], dtype=torch.float32)
        box = torch.cat([box[:, :2] - box[:, 2:] / 2, box[:, 2:] + (box[:, :2] - box[:, 2:] / 2)], -1)
        box = torch.stack([box[:, 0].clamp(0, W - 1), box[:, 1].clamp(0, H - 1), box[:, 2].clamp(0, W - 1),
                           box[:, 3].clamp(0, H - 1)], -1)
        keep = ((box[:, 2] - box[:, 0] + 1) >= 0) & ((box[:, 3] - box[:, 1] + 1) >= 0)
        out.append((box[keep], s[keep], idx[keep] + 1))
    return out


def gdino_visual_features(img, image_sizes, sd):
    """Backbone + input projections of GroundingDINO.forward (groundingdino.py:486-516): Swin-T outputs 1..3 -> 1x1 conv + GroupNorm(32),
    plus one 3x3 stride-2 conv + GroupNorm level from the last Swin output; padding masks interpolated per level.
    Returns (srcs: 4 x [B,256,h,w], masks: 4 x bool [B,h,w])."""
    B, _, Hp, Wp = img.shape
    feats = swin_transformer(img, _sub(sd, "backbone.0."))[1:]
    m = torch.zeros(B, Hp, Wp)
    for b, (h, w) in enumerate(image_sizes):
        m[b, h:, :] = 1
        m[b, :, w:] = 1
    srcs, masks = [], []
    for l, f in enumerate(feats):
        y = F.conv2d(f, sd[f"input_proj.{l}.0.weight"], sd[f"input_proj.{l}.0.bias"])
        srcs.append(F.group_norm(y, 32, sd[f"input_proj.{l}.1.weight"], sd[f"input_proj.{l}.1.bias"]))
        masks.append(F.interpolate(m[None], size=f.shape[-2:]).to(torch.bool)[0])
    y = F.conv2d(feats[-1], sd["input_proj.3.0.weight"], sd["input_proj.3.0.bias"], stride=2, padding=1)
    srcs.append(F.group_norm(y, 32, sd["input_proj.3.1.weight"], sd["input_proj.3.1.bias"]))
    masks.append(F.interpolate(m[None], size=y.shape[-2:]).to(torch.bool)[0])
    return srcs, masks


def gdino_forward(img, image_sizes, input_ids, attention_mask, positive_map, bank, sd, K=5, num_classes=80, num_queries=900,
                  enc_layers=6, dec_layers=6, box_threshold=0.05):
    """GroundingDINO.forward at eval (groundingdino.py:447-662) for a batch sharing one prompt: Swin-T (3 outputs) -> input_proj
    (+ one stride-2 level) -> QuerySelector / flatten_fpn_features -> BertModelWarper(QVBertModel) with the per-category text
    masks -> feat_map -> Transformer -> last-layer class logits / boxes -> detections.  img [B,3,H,W] normalised and padded;
    image_sizes list of (h, w); bank {label: [n,1,256]} with n == K."""
    B = img.shape[0]
    T = input_ids.shape[1]
    srcs, masks = gdino_visual_features(img, image_sizes, sd)
    poss = [position_embedding_sine_hw(mk) for mk in masks]
    labels = [k for k, v in positive_map.items() if len(v) != 0]
    vision = torch.cat([bank[l][:K].flatten(0, 1) for l in labels])[None].expand(B, -1, -1)
    vmask = torch.zeros(1, vision.shape[1], T)
    r = 0
    for l in labels:
        n = bank[l][:K].flatten(0, 1).shape[0]
        vmask[0, r:r + n, positive_map[l]] = 1.0
        r += n
    vmask = vmask.expand(B, -1, -1)
    pooled = torch.cat([F.avg_pool2d(f, 2).flatten(2) for f in srcs], dim=2).permute(0, 2, 1)
    ids = input_ids.expand(B, -1) if input_ids.shape[0] == 1 else input_ids
    am = attention_mask.expand(B, -1) if attention_mask.shape[0] == 1 else attention_mask
    self_mask, pid = gdino_text_masks(ids)
    bsd = _sub(sd, "bert.")
    h = bert_embeddings(ids, bsd, position_ids=pid)
    ext = (1.0 - self_mask[:, None].float()) * -10000.0
    vq = preselect(vision, pooled, bsd, "pre_select.")
    for i in range(12):
        if i >= 6:
            h = gcp_block(h, vq, vmask, bsd, f"encoder.qv_layer.{i - 6}.")
        h = bert_layer(h, ext, bsd, f"encoder.layer.{i}.", 12)
    enc_text = _lin(h, sd, "feat_map")
    tsd = _sub(sd, "transformer.")
    tr = gdino_transformer(srcs, masks, poss, enc_text, am.bool(), pid, self_mask, tsd, num_queries=num_queries,
                           enc_layers=enc_layers, dec_layers=dec_layers, return_all=True)
    hs, refs = tr["hs"], tr["references"]
    boxes = (_mlp(hs[-1], sd, f"bbox_embed.{dec_layers - 1}.", 3) + inverse_sigmoid(refs[-2])).sigmoid()
    logits = contrastive_embed(hs[-1], tr["memory_text"], am.bool())
    dets = gdino_detections(logits, boxes, positive_map, num_classes, image_sizes, box_threshold)
    return {"srcs": srcs, "bert_hidden": h, "encoded_text": enc_text, "memory": tr["memory"], "memory_text": tr["memory_text"],
            "hs": hs, "references": refs, "pred_logits": logits, "pred_boxes": boxes, "detections": dets, "topk": tr["topk"],
            "enc_class": tr["enc_class"]}


# ------------------------------------------------------------------------------------------------------------------
# Training side — the backward oracle of the GCP block is torch.autograd over gcp_block() above
# ------------------------------------------------------------------------------------------------------------------
def token_focal_loss(pred_logits, targets, alpha=0.25, gamma=2.0, text_mask=None):
    """TokenSigmoidFocalLoss.forward(version="binary") = token_sigmoid_binary_focal_loss(...).sum()
    (maskrcnn_benchmark/layers/sigmoid_focal_loss.py:127-162,173-184): elements of masked text tokens are dropped."""
    if text_mask is not None:
        keep = (text_mask > 0).unsqueeze(1).expand_as(pred_logits)
        pred_logits, targets = pred_logits[keep], targets[keep]
    p = torch.sigmoid(pred_logits)
    ce = F.binary_cross_entropy_with_logits(pred_logits, targets, reduction="none")
    p_t = p * targets + (1 - p) * (1 - targets)
    loss = ce * ((1 - p_t) ** gamma)
    if alpha >= 0:
        loss = (alpha * targets + (1 - alpha) * (1 - targets)) * loss
    return loss.sum()
