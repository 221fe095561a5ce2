"""Two-stage query selection of the GroundingDINO transformer, drop-in for the ``two_stage_type == "standard"`` section of
groundingdino_new/models/GroundingDINO/transformer.py:297-318:

    topk_logits   = enc_outputs_class_unselected.max(-1)[0]
    topk_proposals = torch.topk(topk_logits, num_queries, dim=1)[1]
    refpoint_embed_undetach = gather(enc_outputs_coord_unselected, topk_proposals)     (unsigmoid)
    init_box_proposal        = gather(output_proposals, topk_proposals).sigmoid()
    tgt_undetach             = gather(output_memory, topk_proposals)

All on the device, no host synchronisation: ``mqdet_row_max_f32`` over the 256 text-token logits of each of the ~22k encoder
positions, ``mqdet_topk_desc`` (one CTA per image: radix select, ordered tie fill, bitonic sort of the 900 winners) and
``mqdet_gather_rows_f32``.  Ties are resolved towards the lower index (torch.topk leaves them unspecified).
"""
import torch

from ... import ops
from ..._lib import MqdetError


@torch.no_grad()
def select_queries(enc_outputs_class, enc_outputs_coord_unselected, output_proposals, output_memory, num_queries=900,
                   proposals=None):
    """enc_outputs_class [B,Q,T] fp32 (-inf on padded text tokens allowed), enc_outputs_coord_unselected / output_proposals
    [B,Q,4] fp32, output_memory [B,Q,C] fp32 -> dict(topk_proposals int64 [B,k], refpoint_embed [B,k,4], init_box_proposal
    [B,k,4], tgt [B,k,C], topk_logits [B,Q])."""
    for t in (enc_outputs_class, enc_outputs_coord_unselected, output_proposals, output_memory):
        if not t.is_cuda:
            raise MqdetError("select_queries: CUDA tensors required (no CPU fallback)")
    topk_logits = ops.row_max(enc_outputs_class.float().contiguous())
    # ``proposals`` (int64 [B, k]): a given selection instead of the top-k -- parity tests feed the oracle's selection so that the
    # decoder can be compared slot by slot although near-tied class logits may rank differently under fp16 operands
    idx = ops.topk_desc(topk_logits, num_queries) if proposals is None else proposals.to(torch.int64).contiguous()
    ref = ops.gather_rows(enc_outputs_coord_unselected.float().contiguous(), idx)
    prop = ops.gather_rows(output_proposals.float().contiguous(), idx, sigmoid=True)
    tgt = ops.gather_rows(output_memory.float().contiguous(), idx)
    return {"topk_logits": topk_logits, "topk_proposals": idx, "refpoint_embed": ref,
            "init_box_proposal": prop, "tgt": tgt}
