"""Box pooler of the vision-query extraction, drop-in for maskrcnn_benchmark/modeling/poolers.py:11-129 (``LevelMapper``,
``Pooler`` with ``use_v2=True``: torchvision ``roi_align(aligned=True)`` per FPN level).

One kernel (``mqdet_roi_align_levels``) maps every box to its level (FPN paper eq. 1) and pools it from the fp16 NHWC pyramid;
``forward_mean`` additionally folds the ``mean(dim=[-2, -1])`` of ``extract_query`` (generalized_vl_rcnn_new.py:263) into it.
"""
import torch
from torch import nn

from .. import ops
from .._lib import MqdetError


class Pooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio, use_v2=False):
        super().__init__()
        if not use_v2:
            raise NotImplementedError("the MQ-Det detector builds its pooler with use_v2=True (generalized_vl_rcnn_new.py:110-122)")
        self.output_size = tuple(output_size) if isinstance(output_size, (tuple, list)) else (output_size, output_size)
        if self.output_size[0] != self.output_size[1]:
            raise NotImplementedError("square pooling only (POOLER_RESOLUTION)")
        self.scales = tuple(float(s) for s in scales)
        self.sampling_ratio = int(sampling_ratio)

    @staticmethod
    def convert_to_roi_format(boxes):
        """[(K1,4), (K2,4)] BoxLists -> rois [K1+K2, 5] with the image index in column 0 (poolers.py:80-97)."""
        rois = []
        for i, b in enumerate(boxes):
            bb = b.bbox.float()
            rois.append(torch.cat([torch.full((bb.shape[0], 1), float(i), dtype=torch.float32, device=bb.device), bb], dim=1))
        return torch.cat(rois, dim=0) if rois else torch.zeros((0, 5))

    def _levels(self, sizes, device):
        lv = ops.get_levels(sizes, device)
        if lv.n != len(self.scales):
            raise MqdetError(f"pooler has {len(self.scales)} scales but the pyramid has {lv.n} levels")
        return lv

    @torch.no_grad()
    def forward_flat(self, pyr16, levels, boxes, mean_only=False):
        rois = self.convert_to_roi_format(boxes).to(pyr16.device)
        out, lvl = ops.roi_align_levels(pyr16, levels, self.scales, rois, self.output_size[0], self.sampling_ratio, mean_only)
        return out, lvl

    @torch.no_grad()
    def forward(self, x, boxes):
        """Reference signature: x = list of [B,C,h,w] maps, boxes = list[BoxList] -> [R, C, P, P] fp32."""
        if not x[0].is_cuda:
            raise MqdetError("Pooler: CUDA tensors required (no CPU fallback)")
        levels = self._levels([(f.shape[2], f.shape[3]) for f in x], x[0].device)
        pyr16 = ops.cast_f16(torch.cat([f.flatten(2).transpose(1, 2) for f in x], dim=1).contiguous())
        return self.forward_flat(pyr16, levels, boxes, mean_only=False)[0]
