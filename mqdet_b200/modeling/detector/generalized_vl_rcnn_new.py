"""MQ-GLIP detector meta-architecture (inference), drop-in for
maskrcnn_benchmark/modeling/detector/generalized_vl_rcnn_new.py:90-667 (``GeneralizedVLRCNN_New``).

    images -> Swin-T -> FPN -> [B, N, 256] pyramid ---------------------------------------------+
    positive_map -> QuerySelector -> vision queries -> PreSelect(queries x pooled pyramid) ----+ |
    token ids -> BERT layers 0..5 -> [GCP_i -> BERT layer 6+i] x 6 -> text stream [B,T,768] --+-+-> VLDyHead -> ATSS
                                                                                                   post-processing -> BoxList

Parameter names follow the reference (``backbone.body.*``, ``backbone.fpn.*``, ``language_backbone.body.model.*``,
``rpn.head.*``) so its checkpoints load unchanged.  Differences from the reference, all at the host boundary:
  * captions may be given pre-tokenised ({"input_ids", "attention_mask"}) — no tokenizer vocabulary is available
    offline; with an HF tokenizer attached (``self.tokenizer``) plain strings work as in the reference;
  * the eval-time ``assert B == 1`` (:354) is lifted: a batch shares one prompt (the training-branch tensor layout).
"""
from collections import OrderedDict

import torch
from torch import nn

from ... import ops
from ..._lib import MqdetError
from ...structures.image_list import to_image_list
from ..backbone.fpn import FPN, LastLevelP6P7
from ..backbone.swint import SwinTransformer
from ..language_backbone.bert_model_new import BertEncoder
from ..query_selector.query_selector import QuerySelector
from ..rpn.vldyhead import VLDyHeadModule


def build_backbone(cfg):
    """SWINT-FPN-RETINANET (modeling/backbone/__init__.py:37-80)."""
    sw = cfg.MODEL.SWINT
    body = SwinTransformer(embed_dim=sw.EMBED_DIM, depths=tuple(sw.DEPTHS), num_heads=tuple(sw.NUM_HEADS),
                           window_size=sw.WINDOW_SIZE, mlp_ratio=sw.MLP_RATIO)
    oc = cfg.MODEL.BACKBONE.OUT_CHANNELS
    fpn = FPN([0, sw.OUT_CHANNELS[-3], sw.OUT_CHANNELS[-2], sw.OUT_CHANNELS[-1]], oc, top_blocks=LastLevelP6P7(oc, oc))
    return nn.Sequential(OrderedDict([("body", body), ("fpn", fpn)]))


class GeneralizedVLRCNN_New(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.backbone = build_backbone(cfg)
        self.language_backbone = nn.Sequential(OrderedDict([("body", BertEncoder(cfg))]))
        self.rpn = VLDyHeadModule(cfg)
        self.roi_heads = None  # RPN_ONLY: True (configs/pretrain/mq-glip-t.yaml:6)
        self.query_selector = QuerySelector(cfg) if cfg.VISION_QUERY.ENABLED else None
        self.tokenizer = None  # attach an HF tokenizer to accept string captions
        self.DEBUG = False
        self._prompt = None

    def load_query_bank(self, path):
        self.query_selector.load_query_bank(path)
        self.invalidate_prompt_cache()

    def invalidate_prompt_cache(self):
        """Drop everything cached per prompt (token ids, selected queries, masks, the head's token map)."""
        self._prompt = None
        self.rpn._tokmap = None

    @staticmethod
    def _prompt_key(captions, positive_map, B, bank_version):
        """Content key of a prompt: token ids / caption strings, the positive map's entries, batch size and the query
        bank version — an in-place edit of ``positive_map`` or a swapped bank is a different prompt."""
        if isinstance(captions, dict):
            ids = captions["input_ids"]
            am = captions["attention_mask"]
            if ids.is_cuda:  # no device->host sync on the hot path: storage identity + version counters
                ck = (tuple(ids.shape), ids.data_ptr(), ids._version, am.data_ptr(), am._version)
            else:
                ck = (tuple(ids.shape), ids.detach().numpy().tobytes(), am.detach().numpy().tobytes())
        else:
            ck = tuple(captions) if isinstance(captions, (list, tuple)) else captions
        pk = tuple((int(k), tuple(int(t) for t in (v if not isinstance(v, int) else [v]))) for k, v in sorted(positive_map.items()))
        return (ck, pk, int(B), int(bank_version))

    def max_out(self):
        """Rows of the fixed-shape per-image result: DETECTIONS_PER_IMG plus head-room for the reference's `>= kth score`
        ties (rpn/inference.py:757-767), rounded up to a multiple of 32 (100 -> 128, LVIS 300 -> 352)."""
        d = int(self.cfg.MODEL.ATSS.DETECTIONS_PER_IMG)
        return (d + 28 + 31) // 32 * 32

    @torch.no_grad()
    def get_labels_and_maps_from_positive_map(self, positive_map, dtype=torch.float):
        """:295-305 — labels with at least one token and their normalised token-location rows."""
        labels = [k for k, v in positive_map.items() if len(v) != 0]
        T = self.cfg.MODEL.LANGUAGE_BACKBONE.MAX_QUERY_LEN
        all_map = torch.zeros((len(labels), T), dtype=dtype)
        for j, label in enumerate(labels):
            all_map[j, positive_map[label]] = 1
        all_map = all_map / (all_map.sum(-1)[:, None] + 1e-6)
        return labels, all_map

    def _tokenize(self, captions, device):
        if isinstance(captions, dict):
            return captions["input_ids"].to(device), captions["attention_mask"].to(device)
        if self.tokenizer is None:
            raise MqdetError("string captions need a tokenizer (bert-base-uncased vocabulary is not available offline): "
                             "pass {'input_ids', 'attention_mask'} or set model.tokenizer")
        L = self.cfg.MODEL.LANGUAGE_BACKBONE
        tok = self.tokenizer.batch_encode_plus(captions, max_length=L.MAX_QUERY_LEN,
                                               padding="max_length" if L.PAD_MAX else "longest",
                                               return_special_tokens_mask=True, return_tensors="pt", truncation=True)
        return tok.input_ids.to(device), tok.attention_mask.to(device)

    @torch.no_grad()
    def _prompt_state(self, captions, positive_map, B, dev):
        """Everything that depends only on the prompt (token ids, selected vision queries, their token mask) is built
        once per (captions, positive_map) object pair and batch size — the reference rebuilds it in every forward
        (Python loops over classes in QuerySelector.forward :57-100 and a tokenizer call :378-383)."""
        st = self._prompt
        bank_version = self.query_selector.bank_version if self.query_selector is not None else 0
        same_objects = st is not None and st["captions"] is captions and st["positive_map"] is positive_map
        key = self._prompt_key(captions, positive_map, B, bank_version)
        if st is not None and st["key"] == key:
            if not same_objects:
                st["captions"], st["positive_map"] = captions, positive_map
            return st
        self.rpn._tokmap = None
        ids, am = self._tokenize(captions, dev)
        if ids.shape[0] == 1 and B > 1:
            ids, am = ids.expand(B, -1).contiguous(), am.expand(B, -1).contiguous()
        vision = vmask = None
        if self.query_selector is not None and self.query_selector.query_bank is not None:
            labels, all_map = self.get_labels_and_maps_from_positive_map(positive_map)
            vision, vmask, _ = self.query_selector([labels] * B, [all_map] * B, None)
            vision, vmask = vision.float().contiguous(), vmask.float().contiguous()
        self._prompt = dict(key=key, captions=captions, positive_map=positive_map, B=B, ids=ids, am=am, vision=vision,
                            vmask=vmask)
        return self._prompt

    @torch.no_grad()
    def forward_device(self, images, captions, positive_map, max_out=None):
        """Everything up to (and excluding) the device->host copy: returns the device-resident result dict.
        ``max_out`` (rows of the fixed-shape detections) defaults to ``self.max_out()`` (from DETECTIONS_PER_IMG)."""
        max_out = self.max_out() if max_out is None else int(max_out)
        if self.training:
            raise NotImplementedError("training is SURVEY.md §8f")
        images = to_image_list(images, self.cfg.DATALOADER.SIZE_DIVISIBILITY)
        x = images.tensors
        if not x.is_cuda:
            raise MqdetError("GeneralizedVLRCNN_New: CUDA images required (no CPU fallback)")
        B = x.shape[0]
        feats = self.backbone.body.forward_flat(x)
        pyr16, levels = self.backbone.fpn.forward_flat([feats[i] for i in (1, 2, 3)])
        st = self._prompt_state(captions, positive_map, B, x.device)
        pooled = ops.avgpool2_levels(pyr16, levels) if st["vision"] is not None else None  # flatten_fpn_features (:291-293)
        lang = self.language_backbone.body({"input_ids": st["ids"], "attention_mask": st["am"],
                                            "vision_inputs": {"vision": st["vision"], "images": pooled,
                                                              "vision_attention_mask": st["vmask"],
                                                              "batched_pos_category_map": None}})
        out = self.rpn.forward_flat(pyr16, levels, images.image_sizes, lang["hidden"], lang["masks"], positive_map, max_out)
        out["image_sizes"] = images.image_sizes
        out["pyramid16"], out["lang_hidden"] = pyr16, lang["hidden"]  # inputs of the fusion tower (parity tests read them)
        out["vision_query_gates"] = lang["vision_query_gates"]
        return out

    def forward(self, images, targets=None, captions=None, positive_map=None, greenlight_map=None,
                return_backbone_features=False):
        """Reference signature (:307-314), eval: returns list[BoxList] (fields ``labels``, ``scores``; mode xyxy)."""
        out = self.forward_device(images, captions, positive_map)
        return self.rpn.to_boxlists(out["det"], out["num"], out["image_sizes"])
