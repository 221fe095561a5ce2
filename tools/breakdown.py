"""Stage-level time breakdown of one forward (CUDA events around the module fast paths) + top kernels by launch-count
free estimate.  Not a benchmark: it only tells where the step time goes."""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from mqdet_b200 import ops
from mqdet_b200.config import mq_glip_t_cfg
from mqdet_b200.modeling.detector.generalized_vl_rcnn_new import GeneralizedVLRCNN_New
from mqdet_b200.structures.image_list import ImageList
from tools import synth

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
gen, ids, am, pmap, bank, img = bench.build_inputs(B, 1235)
sd = synth.detector_sd(synth.Gen(99), bias0=-4.59511985013459)
model = GeneralizedVLRCNN_New(mq_glip_t_cfg())
for k, v in model.state_dict().items():
    if k.endswith("relative_position_index"):
        sd[k] = v
model.load_state_dict(sd, strict=True)
del sd
model = model.to(dev).eval()
model.query_selector.set_query_bank(bank)
model.rpn.head.overlap_text_stream = os.environ.get("MQDET_OVERLAP", "0") == "1"  # stage times: branches one after the other
caps = {"input_ids": ids, "attention_mask": am}
x = img.to(dev)
sizes = [(bench.H_IMG, bench.W_IMG)] * B
events = []


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        events.append((label, e0, e1))
        return r
    setattr(obj, name, w)


wrap(model.backbone.body, "forward_flat", "swin")
wrap(model.backbone.fpn, "forward_flat", "fpn")
wrap(model.language_backbone.body, "forward", "language (bert+gcp+preselect)")
wrap(model.language_backbone.body.model.pre_select, "forward", "  preselect")
tower = model.rpn.head.dyhead_tower
for i in range(0, len(tower), 3):
    wrap(tower[i].b_attn, "forward_flat", "fusion: biattention")
    wrap(tower[i + 2], "forward_flat", "fusion: dyconv")
wrap(model.rpn.head, "forward_flat", "head total (tower + dot head)")
wrap(ops, "atss_postprocess", "postprocess (atss + nms)")
import mqdet_b200.modeling.language_backbone.modeling_bert_new as mb
orig_bert = mb.BertLayer.forward


def bert_fwd(self, *a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = orig_bert(self, *a, **k)
    e1.record()
    events.append(("bert layers (18)", e0, e1))
    return r


mb.BertLayer.forward = bert_fwd
for _ in range(3):
    events.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    model.forward_device(ImageList(x, sizes), caps, pmap)
    t1.record()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for label, a, b in events:
    agg[label] = agg.get(label, 0.0) + a.elapsed_time(b)
total = t0.elapsed_time(t1)
print(f"total {total:.2f} ms (B={B})")
for k, v in agg.items():
    print(f"  {v:8.2f} ms  {100*v/total:5.1f}%  {k}")
# per C-ABI entry point (all of the product's kernels go through these) + GEMM shapes
from mqdet_b200 import _lib
_lib.CALL_PROFILE = []
ops.GEMM_PROFILE = []
t0.record()
model.forward_device(ImageList(x, sizes), caps, pmap)
t1.record()
torch.cuda.synchronize()
calls, prof = _lib.CALL_PROFILE, ops.GEMM_PROFILE
_lib.CALL_PROFILE = None
ops.GEMM_PROFILE = None
by = {}
for name, a, b in calls:
    e = by.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += a.elapsed_time(b)
tot = sum(v[1] for v in by.values())
print(f"C-ABI calls: {sum(v[0] for v in by.values())} launches, {tot:.2f} ms inside them, step (with event overhead) {t0.elapsed_time(t1):.2f} ms")
for name, (c, ms) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"  {ms:8.2f} ms {100 * ms / tot:5.1f}%  x{c:4d}  {name}")
sh = {}
for a, b, fl, shape in prof:
    e = sh.setdefault(shape, [0, 0.0, 0.0])
    e[0] += 1
    e[1] += a.elapsed_time(b)
    e[2] += fl
print("GEMM shapes (M,N,K,batch): count, ms, TFLOP/s")
for shape, (c, ms, fl) in sorted(sh.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"  {shape}: x{c} {ms:7.2f} ms {fl / ms / 1e9:7.1f}")
json.dump({"total_ms": total, "stages": agg, "calls": {k: v for k, v in by.items()},
           "gemm": {str(k): v for k, v in sh.items()}}, open(os.path.join(ROOT, "gpurun_out", "breakdown.json"), "w"), indent=1)
