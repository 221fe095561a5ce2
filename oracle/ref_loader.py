"""TEST INFRASTRUCTURE ONLY — loads the reference's own hot-path modules from /root/reference, unmodified.

Used by oracle/make_golden.py (to generate tests/golden/*.pt) and by tests/test_oracle_pinning.py (to pin the
restatement in oracle/restate.py).  /root/reference exists only in the build container, never on the GPU box,
so nothing under ``-m gpu``, ``smoke()`` or ``bench.py`` may import this file.

The reference package cannot be imported as a package here (yacs / timm / einops_exts / maskrcnn_benchmark._C are
absent), so individual files are loaded with importlib after installing small shims:
  * ``einops_exts.rearrange_many``                      (pure reshape helper)
  * five doc-decorator names the file imports from ``transformers.models.bert.modeling_bert`` (gone in HF 5.x)
  * a stub ``maskrcnn_benchmark`` package exposing ``utils.torch_dropout`` (the real file) and ``modeling.utils``
  * ``timm.models.layers.DropPath`` (identity at eval), ``to_2tuple``, ``trunc_normal_``
No reference source is copied into this repository.
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get("MQDET_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "maskrcnn_benchmark"))


def _load_file(mod_name, rel_path):
    path = os.path.join(REF, rel_path)
    spec = importlib.util.spec_from_file_location(mod_name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[mod_name] = mod
    spec.loader.exec_module(mod)
    return mod


def _install_shims():
    import torch
    from einops import rearrange

    if "einops_exts" not in sys.modules:
        m = types.ModuleType("einops_exts")
        m.rearrange_many = lambda tensors, pattern, **kw: tuple(rearrange(t, pattern, **kw) for t in tensors)
        sys.modules["einops_exts"] = m

    import transformers.models.bert.modeling_bert as hf_bert

    def _passthrough(*a, **k):
        def deco(fn):
            return fn
        return deco

    for name, val in [("add_start_docstrings_to_model_forward", _passthrough), ("add_code_sample_docstrings", _passthrough),
                      ("BERT_INPUTS_DOCSTRING", "{}"), ("_CHECKPOINT_FOR_DOC", ""), ("_CONFIG_FOR_DOC", "")]:
        if not hasattr(hf_bert, name):
            setattr(hf_bert, name, val)
    for name in ("BaseModelOutputWithPastAndCrossAttentions", "BaseModelOutputWithPoolingAndCrossAttentions"):
        if not hasattr(hf_bert, name):
            import transformers.modeling_outputs as mo
            setattr(hf_bert, name, getattr(mo, name))
    if not hasattr(hf_bert, "logger"):
        import logging
        hf_bert.logger = logging.getLogger("hf_bert")

    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")

        class DropPath(torch.nn.Module):
            def __init__(self, p=0.0):
                super().__init__()
                self.p = p

            def forward(self, x):
                return x

        layers.DropPath = DropPath
        layers.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models = models
        models.layers = layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})

    if "maskrcnn_benchmark" not in sys.modules:
        pkg = types.ModuleType("maskrcnn_benchmark")
        pkg.__path__ = []
        utils = types.ModuleType("maskrcnn_benchmark.utils")
        utils.__path__ = []
        modeling = types.ModuleType("maskrcnn_benchmark.modeling")
        modeling.__path__ = []
        sys.modules.update({"maskrcnn_benchmark": pkg, "maskrcnn_benchmark.utils": utils,
                            "maskrcnn_benchmark.modeling": modeling})
        pkg.utils, pkg.modeling = utils, modeling
        utils.torch_dropout = _load_file("maskrcnn_benchmark.utils.torch_dropout",
                                         "maskrcnn_benchmark/utils/torch_dropout.py")
        modeling.utils = _load_file("maskrcnn_benchmark.modeling.utils", "maskrcnn_benchmark/modeling/utils.py")


_cache = {}


def modeling_bert_new():
    """maskrcnn_benchmark/modeling/language_backbone/modeling_bert_new.py (GCP, PreSelect, QVBert*)."""
    if "mbn" not in _cache:
        _install_shims()
        _cache["mbn"] = _load_file("ref_modeling_bert_new",
                                   "maskrcnn_benchmark/modeling/language_backbone/modeling_bert_new.py")
    return _cache["mbn"]


def fuse_helper():
    """maskrcnn_benchmark/utils/fuse_helper.py (BiMultiHeadAttention, BiAttentionBlockForCheckpoint)."""
    if "fh" not in _cache:
        _install_shims()
        _cache["fh"] = _load_file("ref_fuse_helper", "maskrcnn_benchmark/utils/fuse_helper.py")
    return _cache["fh"]


def rpn_modeling_bert():
    """maskrcnn_benchmark/modeling/rpn/modeling_bert.py (BertAttention/Intermediate/Output with clamps)."""
    if "rmb" not in _cache:
        _install_shims()
        import transformers.modeling_utils as mu
        import transformers.pytorch_utils as pu
        for name in ("find_pruneable_heads_and_indices", "prune_linear_layer", "apply_chunking_to_forward"):
            if not hasattr(mu, name):
                setattr(mu, name, getattr(pu, name, lambda *a, **k: None))
        _cache["rmb"] = _load_file("ref_rpn_modeling_bert", "maskrcnn_benchmark/modeling/rpn/modeling_bert.py")
    return _cache["rmb"]


def swint():
    """maskrcnn_benchmark/modeling/backbone/swint.py (WindowAttention, SwinTransformerBlock, ...)."""
    if "swint" not in _cache:
        _install_shims()
        _cache["swint"] = _load_file("ref_swint", "maskrcnn_benchmark/modeling/backbone/swint.py")
    return _cache["swint"]


def gdino_utils():
    """groundingdino_new/models/GroundingDINO/utils.py (ContrastiveEmbed; the file only needs torch)."""
    if "gdu" not in _cache:
        _cache["gdu"] = _load_file("ref_gdino_utils", "groundingdino_new/models/GroundingDINO/utils.py")
    return _cache["gdu"]


def gdino_ms_deform_attn():
    """groundingdino_new/models/GroundingDINO/ms_deform_attn.py (its CPU path needs torch only; the failed ``_C`` import is
    caught by the file itself)."""
    if "msda" not in _cache:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _cache["msda"] = _load_file("ref_gdino_msda", "groundingdino_new/models/GroundingDINO/ms_deform_attn.py")
    return _cache["msda"]


def gdino_fuse_modules():
    """groundingdino_new/models/GroundingDINO/fuse_modules.py (BiAttentionBlock / BiMultiHeadAttention; torch + the timm
    DropPath shim)."""
    if "gdfuse" not in _cache:
        _install_shims()
        _cache["gdfuse"] = _load_file("ref_gdino_fuse", "groundingdino_new/models/GroundingDINO/fuse_modules.py")
    return _cache["gdfuse"]


def query_selector():
    """maskrcnn_benchmark/modeling/query_selector/query_selector.py (pure torch / numpy)."""
    if "qs" not in _cache:
        _cache["qs"] = _load_file("ref_query_selector", "maskrcnn_benchmark/modeling/query_selector/query_selector.py")
    return _cache["qs"]


def bounding_box():
    """maskrcnn_benchmark/structures/bounding_box.py (BoxList)."""
    if "bb" not in _cache:
        _cache["bb"] = _load_file("ref_bounding_box", "maskrcnn_benchmark/structures/bounding_box.py")
    return _cache["bb"]


def image_list():
    """maskrcnn_benchmark/structures/image_list.py (ImageList, to_image_list)."""
    if "il" not in _cache:
        _cache["il"] = _load_file("ref_image_list", "maskrcnn_benchmark/structures/image_list.py")
    return _cache["il"]


def rpn_inference(ml_nms_fn):
    """maskrcnn_benchmark/modeling/rpn/inference.py (ATSSPostProcessor, convert_grounding_to_od_logits) together with the
    reference's own box_coder.py, bounding_box.py and boxlist_ops.py.  The only thing the file cannot get on CPU is the
    compiled ``maskrcnn_benchmark.layers.ml_nms`` (a CUDA kernel): ``ml_nms_fn(boxes, scores, labels, thresh) -> kept
    indices`` stands in for it (the kernel itself is pinned on the GPU, tests/test_ref_kernels_gpu.py)."""
    _install_shims()
    pkg = sys.modules["maskrcnn_benchmark"]
    if "maskrcnn_benchmark.layers" not in sys.modules:
        layers = types.ModuleType("maskrcnn_benchmark.layers")
        sys.modules["maskrcnn_benchmark.layers"] = layers
        pkg.layers = layers
    layers = sys.modules["maskrcnn_benchmark.layers"]
    layers.ml_nms = ml_nms_fn
    layers.nms = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("single-class nms is not on the MQ-Det path"))
    if "rpn_inf" not in _cache:
        structures = types.ModuleType("maskrcnn_benchmark.structures")
        structures.__path__ = []
        sys.modules["maskrcnn_benchmark.structures"] = structures
        pkg.structures = structures
        structures.bounding_box = _load_file("maskrcnn_benchmark.structures.bounding_box",
                                             "maskrcnn_benchmark/structures/bounding_box.py")
        structures.boxlist_ops = _load_file("maskrcnn_benchmark.structures.boxlist_ops",
                                            "maskrcnn_benchmark/structures/boxlist_ops.py")
        modeling = sys.modules["maskrcnn_benchmark.modeling"]
        modeling.box_coder = _load_file("maskrcnn_benchmark.modeling.box_coder", "maskrcnn_benchmark/modeling/box_coder.py")
        rpn = types.ModuleType("maskrcnn_benchmark.modeling.rpn")
        rpn.__path__ = []
        sys.modules["maskrcnn_benchmark.modeling.rpn"] = rpn
        modeling.rpn = rpn
        _cache["rpn_inf"] = _load_file("maskrcnn_benchmark.modeling.rpn.inference", "maskrcnn_benchmark/modeling/rpn/inference.py")
    inf = _cache["rpn_inf"]

    def boxlist_ml_nms_gpu_branch(boxlist, nms_thresh, max_proposals=-1, score_field="scores", label_field="labels"):
        # boxlist_ops.py:48-75 as executed for CUDA tensors (:69); on CPU tensors the reference takes a per-label debugging
        # branch (:56-67) that is not what inference runs, so the wrapper is re-stated here around the substituted kernel
        if nms_thresh <= 0:
            return boxlist
        mode = boxlist.mode
        boxlist = boxlist.convert("xyxy")
        keep = ml_nms_fn(boxlist.bbox, boxlist.get_field(score_field), boxlist.get_field(label_field).float(), nms_thresh)
        if max_proposals > 0:
            keep = keep[:max_proposals]
        return boxlist[keep].convert(mode)

    inf.boxlist_ml_nms = boxlist_ml_nms_gpu_branch
    return inf


def vldyhead(dcn_fn):
    """maskrcnn_benchmark/modeling/rpn/vldyhead.py (DyConv, BertEncoderLayer, VLFuse, VLDyHead) with the reference's own
    dyrelu.py / se.py / misc.py / fuse_helper.py / rpn/modeling_bert.py / modeling_bert_new.py.  Not loadable pieces are
    stubbed: the loss / anchor factories (never called by VLDyHead.forward), fbnet, clip_model, and the compiled
    ``ModulatedDeformConv`` — a module with the same parameters whose forward calls
    ``dcn_fn(x, offset, mask, weight, bias, stride)`` (the CUDA kernel itself is pinned on the GPU,
    tests/test_ref_kernels_gpu.py)."""
    import torch
    from torch import nn
    if "vldyhead" in _cache:
        _cache["vldyhead"]._dcn_fn[0] = dcn_fn
        return _cache["vldyhead"]
    if "rpn_inf" not in _cache:
        rpn_inference(lambda *a: (_ for _ in ()).throw(NotImplementedError))
    fh = fuse_helper()
    mbn = modeling_bert_new()
    rmb = rpn_modeling_bert()
    pkg = sys.modules["maskrcnn_benchmark"]
    modeling = sys.modules["maskrcnn_benchmark.modeling"]
    utils = sys.modules["maskrcnn_benchmark.utils"]
    sys.modules["maskrcnn_benchmark.utils.fuse_helper"] = fh
    utils.fuse_helper = fh
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    if not hasattr(mu, "apply_chunking_to_forward"):
        mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    dcn_holder = [dcn_fn]

    class ModulatedDeformConv(nn.Module):  # parameters / signature of layers/deform_conv.py:335-384, kernel substituted
        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                     deformable_groups=1, bias=True):
            super().__init__()
            assert kernel_size == 3 and padding == 1 and dilation == 1 and groups == 1 and deformable_groups == 1
            self.stride = stride
            self.weight = nn.Parameter(torch.zeros(out_channels, in_channels, 3, 3))
            self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

        def forward(self, input, offset, mask):
            return dcn_holder[0](input, offset, mask, self.weight, self.bias, self.stride)

    layers = sys.modules["maskrcnn_benchmark.layers"]
    layers.__path__ = []
    dy = _load_file("maskrcnn_benchmark.layers.dyrelu", "maskrcnn_benchmark/layers/dyrelu.py")
    se = _load_file("maskrcnn_benchmark.layers.se", "maskrcnn_benchmark/layers/se.py")
    misc = _load_file("maskrcnn_benchmark.layers.misc", "maskrcnn_benchmark/layers/misc.py")
    layers.DYReLU, layers.SELayer, layers.Scale = dy.DYReLU, se.SELayer, misc.Scale
    layers.ModulatedDeformConv = ModulatedDeformConv
    layers.NaiveSyncBatchNorm2d = layers.FrozenBatchNorm2d = nn.BatchNorm2d  # bn types the MQ configs never select

    def stub(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod
        return mod

    unused = lambda *a, **k: None  # noqa: E731
    stub("maskrcnn_benchmark.modeling.rpn.loss", make_atss_loss_evaluator=unused)
    stub("maskrcnn_benchmark.modeling.rpn.anchor_generator", make_anchor_generator_complex=unused)
    stub("maskrcnn_benchmark.modeling.backbone", __path__=[])
    import math as _math
    stub("maskrcnn_benchmark.modeling.backbone.fbnet", math=_math)  # vldyhead.py gets `math` through this star import
    stub("maskrcnn_benchmark.engine", __path__=[])
    stub("maskrcnn_benchmark.engine.inference", create_positive_map_label_to_token_from_positive_map=unused)
    stub("maskrcnn_benchmark.modeling.language_backbone", __path__=[])
    stub("maskrcnn_benchmark.modeling.language_backbone.clip_model", QuickGELU=nn.GELU, LayerNorm=nn.LayerNorm,
         DropPath=sys.modules["timm.models.layers"].DropPath)
    sys.modules["maskrcnn_benchmark.modeling.language_backbone.modeling_bert_new"] = mbn
    sys.modules["maskrcnn_benchmark.modeling.rpn.modeling_bert"] = rmb
    mod = _load_file("maskrcnn_benchmark.modeling.rpn.vldyhead", "maskrcnn_benchmark/modeling/rpn/vldyhead.py")
    # environment adaptations (no arithmetic of the head is touched):
    from transformers import BertConfig

    class _OfflineBertConfig(BertConfig):  # VLDyHead.__init__ :600 asks the hub; bert-base-uncased == BertConfig()
        @classmethod
        def from_pretrained(cls, name, **k):
            assert name == "bert-base-uncased"
            return BertConfig()

    mod.BertConfig = _OfflineBertConfig
    # transformers-4 signature (mask, input_shape, device) and additive value of the PreTrainedModel helper that
    # BertEncoderLayer.forward calls (:273); transformers 5 takes a dtype in third position.  -10000 vs finfo.min is
    # immaterial: exp() underflows to exactly 0 either way.
    mod.BertEncoderLayer.get_extended_attention_mask = \
        lambda self, mask, shape, device=None: (1.0 - mask[:, None, None, :].float()) * -10000.0
    mod._dcn_fn = dcn_holder
    _cache["vldyhead"] = mod
    return mod


def anchor_generator():
    """maskrcnn_benchmark/modeling/rpn/anchor_generator.py (AnchorGenerator, make_anchor_generator_complex) on the
    reference's own BoxList / ImageList."""
    if "ag" not in _cache:
        if "rpn_inf" not in _cache:
            rpn_inference(lambda *a: None)  # installs the structures package (bounding_box, boxlist_ops)
        st = sys.modules["maskrcnn_benchmark.structures"]
        st.image_list = _load_file("maskrcnn_benchmark.structures.image_list", "maskrcnn_benchmark/structures/image_list.py")
        _cache["ag"] = _load_file("ref_anchor_generator", "maskrcnn_benchmark/modeling/rpn/anchor_generator.py")
    return _cache["ag"]


def detector(cfg, dcn_fn, ml_nms_fn, tokenized):
    """An instance of the reference's ``GeneralizedVLRCNN_New`` (modeling/detector/generalized_vl_rcnn_new.py) assembled from
    the reference's own parts on CPU: Swin-T + FPN, QuerySelector, ``BertEncoder`` -> ``QVBertModel``, ``VLDyHeadModule``
    (VLDyHead + AnchorGenerator + ATSSPostProcessor).  ``__init__`` is bypassed (it needs yacs registries, the HF hub and
    the compiled ``_C``); the attributes it would create are set here with the same classes and the mq-glip-t configuration.
    Substitutions: the two compiled kernels (``dcn_fn``, ``ml_nms_fn``), the transformers-4 shims of ``vldyhead`` /
    ``QVBertModel`` (see there), and the tokenizer (no vocabulary offline): ``tokenized`` = (input_ids, attention_mask)
    is what ``batch_encode_plus`` returns for any caption."""
    import torch
    from collections import OrderedDict
    from torch import nn
    from transformers import BertConfig
    vd = vldyhead(dcn_fn)
    ag = anchor_generator()
    rpn_inference(ml_nms_fn)  # last: (re)binds the NMS substitute
    vd.make_anchor_generator_complex = ag.make_anchor_generator_complex
    mbn, rmb = modeling_bert_new(), rpn_modeling_bert()
    sw, fp, qs = swint(), fpn(), query_selector()
    modeling = sys.modules["maskrcnn_benchmark.modeling"]

    def stub(name, **attrs):
        mod = sys.modules.get(name) or types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod
        return mod

    unused = lambda *a, **k: None  # noqa: E731
    if "det" not in _cache:
        stub("maskrcnn_benchmark.modeling.poolers", CustomPooler=object, Pooler=object)
        stub("maskrcnn_benchmark.modeling.backbone", build_backbone=unused)
        stub("maskrcnn_benchmark.modeling.rpn", build_rpn=unused)
        stub("maskrcnn_benchmark.modeling.roi_heads", build_roi_heads=unused)
        stub("maskrcnn_benchmark.modeling.query_selector", build_query_selector=unused)
        stub("maskrcnn_benchmark.modeling.language_backbone", build_language_backbone=unused)
        stub("maskrcnn_benchmark.modeling.detector", __path__=[])
        _cache["bmn"] = _load_file("maskrcnn_benchmark.modeling.language_backbone.bert_model_new",
                                   "maskrcnn_benchmark/modeling/language_backbone/bert_model_new.py")
        _cache["det"] = _load_file("maskrcnn_benchmark.modeling.detector.generalized_vl_rcnn_new",
                                   "maskrcnn_benchmark/modeling/detector/generalized_vl_rcnn_new.py")
    det_mod, bmn = _cache["det"], _cache["bmn"]
    config = BertConfig()

    # ---- language backbone: BertEncoder wrapper around QVBertModel (12 layer adapters, see tests) ----
    qv = mbn.QVBertModel(config, dim_t=768, dim_v=256, cfg=cfg, add_pooling_layer=False)
    qv.encoder.gradient_checkpointing = False

    class Layer(nn.Module):  # positional interface of the transformers-4 BertLayer, built from the reference's in-repo copy
        def __init__(self):
            super().__init__()
            self.attention = rmb.BertAttention(config, False, False)
            self.intermediate = rmb.BertIntermediate(config)
            self.output = rmb.BertOutput(config)

        def forward(self, h, attention_mask=None, head_mask=None, enc_h=None, enc_mask=None, past=None, output_attentions=False):
            a = self.attention(h, attention_mask, None, output_attentions=False, past_key_value=None)[0]
            return (self.output(self.intermediate(a), a),)

    qv.encoder.layer = nn.ModuleList([Layer() for _ in range(config.num_hidden_layers)])
    if not hasattr(qv.embeddings, "position_embedding_type"):
        qv.embeddings.position_embedding_type = "absolute"
    if not hasattr(qv, "get_head_mask"):
        qv.get_head_mask = lambda head_mask, n, *a, **k: [None] * n
    body = bmn.BertEncoder.__new__(bmn.BertEncoder)
    nn.Module.__init__(body)
    body.cfg, body.bert_name, body.model, body.language_dim, body.num_layers = cfg, "bert-base-uncased", qv, 768, 1
    lang = nn.Sequential(OrderedDict([("body", body)]))

    # ---- visual backbone: Swin-T -> FPN (+P6, P7), as backbone/__init__.py:37-80 assembles it ----
    body_v = sw.SwinTransformer(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7, drop_path_rate=0.0,
                                frozen_stages=-1, use_checkpoint=False)
    conv_block = lambda i, o, k, s=1: nn.Conv2d(i, o, k, s, padding=(k - 1) // 2)  # noqa: E731
    fpn_v = fp.FPN([0, 192, 384, 768], 256, conv_block, top_blocks=fp.LastLevelP6P7(256, 256))
    backbone = nn.Sequential(OrderedDict([("body", body_v), ("fpn", fpn_v)]))

    # ---- head module ----
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        rpn = vd.VLDyHeadModule(cfg)

    class Tok:  # what batch_encode_plus(...) returns, for any caption
        input_ids, attention_mask = tokenized
        special_tokens_mask = 1 - tokenized[1]

        def to(self, device):
            return self

    class Tokenizer:
        mask_token_id, pad_token_id = 103, 0

        def batch_encode_plus(self, captions, **kw):
            return Tok()

    det = det_mod.GeneralizedVLRCNN_New.__new__(det_mod.GeneralizedVLRCNN_New)
    nn.Module.__init__(det)
    det.cfg, det.backbone, det.language_backbone, det.rpn, det.roi_heads = cfg, backbone, lang, rpn, None
    det.query_selector = qs.QuerySelector(cfg)
    det.tokenizer = Tokenizer()
    det.pool = nn.AvgPool2d(2)
    det.use_mlm_loss, det.mlm_loss_for_only_positives, det.force_boxes = False, False, False
    for flag in ("freeze_backbone", "freeze_fpn", "freeze_rpn", "linear_prob", "freeze_cls_logits", "add_linear_layer",
                 "freeze_language_backbone"):  # read by the train()/eval() override (:188-229); all off at inference
        setattr(det, flag, False)
    det.eval()  # the reference's train() override returns None: no chaining
    return det


def poolers():
    """maskrcnn_benchmark/modeling/poolers.py (LevelMapper, Pooler) with the reference's own layers/roi_align.py; the compiled
    ``_C`` is only reached by the non-V2 ``ROIAlign`` (never built by the detector: use_v2=True), so an empty stand-in does."""
    _install_shims()
    if "poolers" not in _cache:
        pkg = sys.modules["maskrcnn_benchmark"]
        if not hasattr(pkg, "_C"):
            pkg._C = types.ModuleType("maskrcnn_benchmark._C")
            sys.modules["maskrcnn_benchmark._C"] = pkg._C
        ra = _load_file("maskrcnn_benchmark.layers.roi_align", "maskrcnn_benchmark/layers/roi_align.py")
        layers = sys.modules.get("maskrcnn_benchmark.layers")
        if layers is None:
            layers = types.ModuleType("maskrcnn_benchmark.layers")
            layers.__path__ = []
            sys.modules["maskrcnn_benchmark.layers"] = layers
            pkg.layers = layers
        layers.ROIAlign, layers.ROIAlignV2 = ra.ROIAlign, ra.ROIAlignV2
        _cache["poolers"] = _load_file("maskrcnn_benchmark.modeling.poolers", "maskrcnn_benchmark/modeling/poolers.py")
    return _cache["poolers"]


def fpn():
    """maskrcnn_benchmark/modeling/backbone/fpn.py (FPN, LastLevelP6P7)."""
    if "fpn" not in _cache:
        _install_shims()
        _cache["fpn"] = _load_file("ref_fpn", "maskrcnn_benchmark/modeling/backbone/fpn.py")
    return _cache["fpn"]


def gdino_package():
    """groundingdino_new/models/GroundingDINO as a pseudo-package ``ref_gdino_pkg`` so that the relative imports of
    ``transformer.py`` (.fuse_modules, .ms_deform_attn, .transformer_vanilla, .utils) resolve to the reference's own files;
    ``groundingdino_new.util.misc`` is the reference's own file loaded by path (only ``inverse_sigmoid`` is used)."""
    if "gdpkg" not in _cache:
        import warnings
        _install_shims()
        if "groundingdino_new" not in sys.modules:
            top = types.ModuleType("groundingdino_new")
            top.__path__ = []
            util = types.ModuleType("groundingdino_new.util")
            util.__path__ = []
            sys.modules.update({"groundingdino_new": top, "groundingdino_new.util": util})
            top.util = util
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                util.misc = _load_file("groundingdino_new.util.misc", "groundingdino_new/util/misc.py")
        pkg = types.ModuleType("ref_gdino_pkg")
        pkg.__path__ = [os.path.join(REF, "groundingdino_new", "models", "GroundingDINO")]
        sys.modules["ref_gdino_pkg"] = pkg
        import importlib
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for sub in ("utils", "fuse_modules", "ms_deform_attn", "transformer_vanilla", "transformer"):
                setattr(pkg, sub, importlib.import_module("ref_gdino_pkg." + sub))
        _cache["gdpkg"] = pkg
    return _cache["gdpkg"]
