"""Diagnostic: which reference detections of the 300-detection case have no counterpart, and why."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from mqdet_b200.config import mq_glip_t_cfg
from mqdet_b200.modeling.detector.generalized_vl_rcnn_new import GeneralizedVLRCNN_New
from mqdet_b200.structures.image_list import ImageList
from oracle import restate, synth
from util import load_sd


def iou_(a, b):
    x1, y1 = torch.max(a[:, None, 0], b[None, :, 0]), torch.max(a[:, None, 1], b[None, :, 1])
    x2, y2 = torch.min(a[:, None, 2], b[None, :, 2]), torch.min(a[:, None, 3], b[None, :, 3])
    inter = (x2 - x1 + 1).clamp(min=0) * (y2 - y1 + 1).clamp(min=0)
    aa = (a[:, 2] - a[:, 0] + 1) * (a[:, 3] - a[:, 1] + 1)
    ab = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
    return inter / (aa[:, None] + ab[None] - inter)


dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=200, precision=4, sci_mode=False)
gen = synth.Gen(2025)
sd = synth.detector_sd(gen, bias0=-1.0)
ids, am, pmap = synth.prompt(10, 2, 256, gen)
bank = synth.query_bank(pmap, 5, gen)
B, h, w = 2, 160, 224
img = synth.images(gen, B, h, w)
print("torch threads", torch.get_num_threads())
ref = restate.detector(img, (h, w), ids, am, pmap, bank, sd, max_det=300)
ref_big = restate.detector(img, (h, w), ids, am, pmap, bank, sd, max_det=100000)
model = GeneralizedVLRCNN_New(mq_glip_t_cfg(**{"MODEL.ATSS.DETECTIONS_PER_IMG": 300}))
full = dict(sd)
for k, v in model.state_dict().items():
    if k.endswith("relative_position_index"):
        full[k] = v
model = load_sd(model, full).to(dev).eval()
model.query_selector.set_query_bank(bank)
il = ImageList(img.to(dev), [(h, w)] * B)
caps = {"input_ids": ids, "attention_mask": am}
res = model(il, captions=caps, positive_map=pmap)
for b in range(B):
    rb, rs, rl = ref["detections"][b]
    ob, os_, ol = res[b].bbox.cpu(), res[b].get_field("scores").cpu(), res[b].get_field("labels").cpu()
    iou = iou_(rb, ob)
    same = rl[:, None] == ol[None]
    m = ((iou > 0.9) & same).any(1)
    print(f"image {b}: ref n={rb.shape[0]} score range [{rs.min():.4f}, {rs.max():.4f}] (all-survivors n={ref_big['detections'][b][0].shape[0]});"
          f" ours n={ob.shape[0]} score range [{os_.min():.4f}, {os_.max():.4f}]; matched {m.float().mean():.3f}")
    print("  ref sorted by score? ", bool((rs[:-1] >= rs[1:]).all()), " ours sorted?", bool((os_[:-1] >= os_[1:]).all()))
    un = (~m).nonzero().flatten()
    print("  unmatched ref idx:", un.tolist()[:80])
    if un.numel():
        bi, bj = iou[un].max(1)
        print("  unmatched: ref score / label / best IoU any label / that box's label / score")
        for k in range(min(12, un.numel())):
            i = un[k].item()
            print(f"    {rs[i]:.4f} {rl[i].item()} {bi[k]:.3f} {ol[bj[k]].item()} {os_[bj[k]]:.4f}")
    # the other way round
    m2 = ((iou > 0.9) & same).any(0)
    un2 = (~m2).nonzero().flatten()
    print(f"  ours without a reference counterpart: {un2.numel()}; their scores: {os_[un2][:12]}")
    # against the un-truncated reference
    rb2, rs2, rl2 = ref_big["detections"][b]
    iou2 = iou_(rb2, ob)
    m3 = ((iou2 > 0.9) & (rl2[:, None] == ol[None])).any(0)
    print(f"  ours found among ALL reference survivors: {m3.float().mean():.3f}")
