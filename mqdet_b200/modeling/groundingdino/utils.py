"""GroundingDINO pieces on the MQ-Det hot path.

Drop-in for ``groundingdino_new/models/GroundingDINO/utils.py:233-268`` (``ContrastiveEmbed``): the query x token
similarity that scores the 900 decoder queries of every decoder layer and the encoder memory tokens of the two-stage
proposal selection (``transformer.py:288-303``).  The deformable encoder/decoder around it is §8(f) "next".
"""
import torch
from torch import nn

from ... import ops
from ..._lib import MqdetError


class ContrastiveEmbed(nn.Module):
    def __init__(self, max_text_len=256):
        super().__init__()
        self.max_text_len = max_text_len

    @torch.no_grad()
    def forward(self, x, text_dict):
        """x [bs, nq, d_model]; text_dict = {'encoded_text': [bs, T, d_model], 'text_token_mask': [bs, T] bool (True =
        token in use)} -> [bs, nq, max_text_len] fp32, -inf on padding and beyond T.  One tcgen05 product (fp16 operands,
        fp32 accumulate/output) + the in-place mask/pad kernel."""
        assert isinstance(text_dict, dict)
        y = text_dict["encoded_text"]
        mask = text_dict["text_token_mask"]
        if not x.is_cuda:
            raise MqdetError("ContrastiveEmbed: CUDA tensors required (no CPU fallback)")
        lead = x.shape[:-2]
        x3 = x.reshape(-1, x.shape[-2], x.shape[-1])
        if y.shape[-1] % 8:
            raise MqdetError("ContrastiveEmbed: d_model must be a multiple of 8")
        out = ops.contrastive_embed(ops.cast_f16(x3.contiguous()), ops.cast_f16(y.contiguous()), mask, self.max_text_len)
        return out.reshape(*lead, x.shape[-2], self.max_text_len)
