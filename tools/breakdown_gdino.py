"""Where the MQ-GroundingDINO-T step time goes (BASELINE config 4, 2 images): CUDA events per stage, per C-ABI entry point and per
GEMM shape of one eager forward.  Not a benchmark.    python tools/breakdown_gdino.py [B]"""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from mqdet_b200 import _lib, ops
from mqdet_b200.config import mq_groundingdino_t_cfg
from mqdet_b200.modeling.groundingdino.groundingdino import GroundingDINO
from mqdet_b200.structures.image_list import ImageList
from tools import synth

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
gen = synth.Gen(1238)
ids, am, pmap = synth.prompt(13, 2, 256, gen)
bank = synth.query_bank(pmap, 5, gen)
img = synth.rgb_images(gen, B, bench.H_IMG, bench.W_IMG)
sd = synth.gdino_sd(synth.Gen(99))
model = GroundingDINO(mq_groundingdino_t_cfg())
for k, v in model.state_dict().items():
    if k.endswith("relative_position_index"):
        sd[k] = v
model.load_state_dict(sd, strict=True)
del sd
model = model.to(dev).eval()
model.query_selector.set_query_bank(bank)
caps = {"input_ids": ids, "attention_mask": am}
il = ImageList(img.to(dev), [(bench.H_IMG, bench.W_IMG)] * B)
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


wrap(model, "visual_features", "swin + input_proj")
wrap(model.bert, "forward", "bert + gcp + preselect")
enc = model.transformer.encoder
for i in range(len(enc.layers)):
    wrap(enc.fusion_layers[i], "forward", "encoder: biattention fusion x6")
    wrap(enc.text_layers[i], "forward_flat", "encoder: text enhancer x6")
    wrap(enc.layers[i], "forward_flat", "encoder: deformable layer x6")
wrap(model.transformer.encoder, "forward_flat", "encoder total")
wrap(model.transformer.decoder, "forward_flat", "decoder total (6 layers)")
wrap(model.transformer, "forward_flat", "transformer total")
for _ in range(3):
    events.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    model.forward_device(il, caps, pmap)
    t1.record()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for label, a, b in events:
    agg[label] = agg.get(label, 0.0) + a.elapsed_time(b)
total = t0.elapsed_time(t1)
print(f"total (eager) {total:.2f} ms (B={B})")
for k, v in agg.items():
    print(f"  {v:8.2f} ms  {100 * v / total:5.1f}%  {k}")
_lib.CALL_PROFILE = []
ops.GEMM_PROFILE = []
t0.record()
model.forward_device(il, caps, pmap)
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
print(f"C-ABI calls: {sum(v[0] for v in by.values())} launches, {tot:.2f} ms inside them")
for name, (c, ms) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"  {ms:8.2f} ms {100 * ms / tot:5.1f}%  x{c:4d}  {name}")
sh = {}
for a, b, fl, shape in prof:
    e = sh.setdefault(shape, [0, 0.0, 0.0])
    e[0] += 1
    e[1] += a.elapsed_time(b)
    e[2] += fl
print("GEMM shapes (M,N,K,batch): count, ms, TFLOP/s")
for shape, (c, ms, fl) in sorted(sh.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"  {shape}: x{c} {ms:7.2f} ms {fl / ms / 1e9:7.1f}")
json.dump({"total_ms": total, "stages": agg, "calls": by, "gemm": {str(k): v for k, v in sh.items()}},
          open(os.path.join(ROOT, "gpurun_out", "breakdown_gdino.json"), "w"), indent=1)
