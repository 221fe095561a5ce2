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


@pytest.mark.parametrize("ref_dim,Q", [(2, 1750), (4, 900)])
def test_ms_deform_attn_vs_oracle(dev, ref_dim, Q):
    """MultiScaleDeformableAttention: encoder-style call (queries == the flattened pyramid, 2-d
    reference points) and decoder-style call (900 queries, 4-d reference boxes), padding mask, batch_first False and True."""
    from mqdet_b200.modeling.groundingdino.ms_deform_attn import MultiScaleDeformableAttention
    from oracle import restate, synth
    from util import assert_close, load_sd
    gen = synth.Gen(1310 + ref_dim)
    sd = synth.msda_sd(gen)
    shapes = [(40, 30), (20, 15), (10, 8), (5, 4)]
    B, E = 2, 256
    nv = sum(h * w for h, w in shapes)
    if ref_dim == 2:
        Q = nv
    query, value = gen.randn(B, Q, E), gen.randn(B, nv, E)
    pos = gen.randn(B, Q, E, scale=0.3)
    ref_pts = torch.rand(B, Q, 4, ref_dim, generator=gen.g)
    if ref_dim == 4:
        ref_pts[..., 2:] = ref_pts[..., 2:] * 0.3 + 0.05
    mask = torch.zeros(B, nv, dtype=torch.bool)
    mask[1, -77:] = True
    ref = restate.ms_deform_attn(query, value, ref_pts, shapes, sd, key_padding_mask=mask, query_pos=pos)
    ss = torch.tensor(shapes)
    lsi = torch.cat([ss.new_zeros(1), (ss[:, 0] * ss[:, 1]).cumsum(0)[:-1]])
    for batch_first in (True, False):
        mod = load_sd(MultiScaleDeformableAttention(embed_dim=E, num_heads=8, num_levels=4, num_points=4, batch_first=batch_first), sd)
        mod = mod.to(dev).eval()
        q, v, p = query.to(dev), value.to(dev), pos.to(dev)
        if not batch_first:
            q, v, p = q.transpose(0, 1), v.transpose(0, 1), p.transpose(0, 1)
        out = mod(q, value=v, query_pos=p, key_padding_mask=mask.to(dev), reference_points=ref_pts.to(dev), spatial_shapes=ss.to(dev),
                  level_start_index=lsi.to(dev))
        if not batch_first:
            out = out.transpose(0, 1)
        assert_close(out, ref, 2e-3, f"MultiScaleDeformableAttention ref_dim={ref_dim} batch_first={batch_first}")
