"""CUDA-event time of the NATIVE part of the modulated pre-training step at the per-GPU shape of BASELINE config 5 (8 images, 80-class
prompt, K = 5 -> 400 vision queries, 5577 pooled image tokens): language-backbone training forward, backward to all 119 trainable
tensors (45.66 M parameters) given dL/d(hidden), global-norm clipping + AdamW.  NOT the whole config-5 step: the frozen fusion tower's
forward / backward and the detection losses are not part of it, so this is a component timing, not a bench line.

    python tools/bench_train_lang.py [B] [steps]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from mqdet_b200 import ops
from mqdet_b200.config import mq_glip_t_cfg
from mqdet_b200.engine.trainer import LanguageSideTrainer
from mqdet_b200.modeling.language_backbone.bert_model_new import bert_base_config
from mqdet_b200.modeling.language_backbone.modeling_bert_new import QVBertModel
from tools import synth

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
gen = synth.Gen(55)
sd = synth.qvbert_sd(gen)
cfg = mq_glip_t_cfg()
model = QVBertModel(bert_base_config(), dim_t=768, dim_v=256, cfg=cfg)
model.load_state_dict(sd, strict=True)
model = model.to(dev)
ids, am, pmap = synth.prompt(80, 2, 256, gen)
ids, am = ids.expand(B, -1).contiguous().to(dev), am.expand(B, -1).contiguous().to(dev)
vision, vmask = synth.vision_queries(pmap, 5, 256, 256, gen)
vision, vmask = vision.expand(B, -1, -1).contiguous().to(dev), vmask.expand(B, -1, -1).contiguous().to(dev)
images = gen.randn(B, 5577, 256).to(dev)
tr = LanguageSideTrainer(model, lr=1e-5, weight_decay=1e-4, clip_value=1.0)
dh = (gen.randn(B, 256, 768) * 1e-2).to(dev)
ev = []
for i in range(2 + steps):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    hid = tr.forward(ids, am, vision, images, vmask)
    e[1].record()
    coef = tr.step(dh)
    e[2].record()
    torch.cuda.synchronize()
    if i >= 2:
        ev.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
fwd = sorted(t[0] for t in ev)[len(ev) // 2]
bwd = sorted(t[1] for t in ev)[len(ev) // 2]
n_par = sum(st["p"].numel() for st in tr.optimizer.params.values())
res = {"what": "native language-side training step (forward + backward of all trainable tensors + clip + AdamW), eager launches",
       "shape": f"B={B}, T=256, 80 classes x 5 queries, 5577 image tokens", "trainable_parameters": n_par,
       "trainable_tensors": len(tr.optimizer.params), "forward_ms": round(fwd, 3), "backward_plus_optimizer_ms": round(bwd, 3),
       "grad_norm": float(coef[1]), "hidden_finite": bool(torch.isfinite(hid).all()), "launches_per_step": ops.launch_count // (2 + steps)}
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "train_lang_step.json"), "w"), indent=1)
