"""GPU parity of the GroundingDINO reuse row: ContrastiveEmbed through the C ABI (tcgen05 product +
mask/pad kernel) against the CPU oracle and the golden vectors recorded from the reference's own class."""
import os

import pytest
import torch

from util import FP16_TOL, ROOT

pytestmark = pytest.mark.gpu


def _check(out, ref, absmax):
    out = out.cpu()
    assert out.shape == ref.shape
    assert torch.equal(torch.isinf(out) & (out < 0), torch.isinf(ref))      # exact -inf pattern (index work: bit-exact)
    fin = torch.isfinite(ref)
    assert (out[fin] - ref[fin]).abs().max().item() <= FP16_TOL * absmax + FP16_TOL


def test_contrastive_embed_vs_oracle_and_golden(dev):
    from mqdet_b200.modeling.groundingdino.utils import ContrastiveEmbed
    from oracle import make_golden, restate
    c = make_golden.case_inputs("contrastive_embed")
    ref = restate.contrastive_embed(c["x"], c["y"], c["mask"], 256)
    mod = ContrastiveEmbed(max_text_len=256)
    out = mod(c["x"].to(dev), {"encoded_text": c["y"].to(dev), "text_token_mask": c["mask"].to(dev)})
    _check(out, ref, ref[torch.isfinite(ref)].abs().max().item())
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "contrastive_embed.pt"))
    _check(make_golden.sub(out.cpu(), *fx["subsample"]["logits"]), fx["logits"], fx["logits_absmax"])


def test_contrastive_embed_encoder_memory_shape(dev):
    """Two-stage proposal scoring (transformer.py:288-303): 22323 memory tokens x 195 text tokens, batch 2."""
    from mqdet_b200.modeling.groundingdino.utils import ContrastiveEmbed
    from oracle import restate
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 22323, 256, generator=g)
    y = torch.randn(2, 195, 256, generator=g)
    mask = torch.ones(2, 195, dtype=torch.bool)
    mask[1, 40:] = False
    ref = restate.contrastive_embed(x, y, mask, 256)
    out = ContrastiveEmbed(256)(x.to(dev), {"encoded_text": y.to(dev), "text_token_mask": mask.to(dev)})
    _check(out, ref, ref[torch.isfinite(ref)].abs().max().item())
    # the two-stage selection consumes the row maximum: same top-900 set up to fp16 ties
    top_ref = ref.max(-1)[0].topk(900, dim=1)[1]
    top_out = out.cpu().max(-1)[0].topk(900, dim=1)[1]
    for b in range(2):
        inter = len(set(top_ref[b].tolist()) & set(top_out[b].tolist()))
        assert inter >= 890, inter
