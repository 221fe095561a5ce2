"""Shared helpers for the parity tests."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# north_star: "outputs match the reference's own forward on identical inputs within 1e-3 fp16 tolerance".
# Read as: max |out - ref| <= 1e-3 * max|ref| + 1e-3 per compared tensor (fp16 operands, fp32 accumulation/statistics).
FP16_TOL = 1e-3


def rel_err(out, ref):
    out = out.detach().float().cpu()
    ref = ref.detach().float().cpu()
    return (out - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)


_records = []


def _dump_records():
    if _records and os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        import json
        with open(os.path.join(ROOT, "gpurun_out", "parity_errors.json"), "w") as f:
            json.dump(_records, f, indent=1)


import atexit  # noqa: E402

atexit.register(_dump_records)


def assert_close(out, ref, tol=FP16_TOL, what="", defer=None):
    out = out.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert out.shape == ref.shape, f"{what}: shape {tuple(out.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(out).all(), f"{what}: non-finite output"
    err = (out - ref).abs().max().item()
    bound = tol * ref.abs().max().item() + tol
    _records.append(dict(what=what, max_abs_err=err, bound=bound, ref_absmax=ref.abs().max().item(),
                         rel=err / (ref.abs().max().item() + 1e-12), mean_abs_err=(out - ref).abs().mean().item(),
                         ref_absmean=ref.abs().mean().item()))
    if defer is not None:
        if err > bound:
            defer.append(f"{what}: max|err| {err:.3e} > {bound:.3e}")
        return err
    assert err <= bound, f"{what}: max|err| {err:.3e} > {bound:.3e} (max|ref| {ref.abs().max().item():.3e})"
    return err


def vq_cfg(**over):
    """The VISION_QUERY block of configs/pretrain/mq-glip-t.yaml:132-141 as an attribute namespace (yacs is absent)."""
    vq = dict(ENABLED=True, FIX_ATTN_GATE=-1.0, CONDITION_GATE=True, NONLINEAR_GATE=True, NO_CAT=True,
              ADD_ADAPT_LAYER=False, RETURN_ATTN_GATE_VALUE=False, VISION_SCALE=1.0, AUGMENT_IMAGE_WITH_QUERY=False,
              TEXT_DROPOUT=0.4, NEW_MASK_TOKEN=False, QUERY_FUSION=False, SHARE_KV=False, NUM_QUERY_PER_CLASS=5)
    vq.update(over)
    return types.SimpleNamespace(VISION_QUERY=types.SimpleNamespace(**vq))


def load_sd(module, sd, prefix=""):
    """Load a flat {name: tensor} dict (reference state_dict names) into a module, strictly."""
    own = module.state_dict()
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    missing = [k for k in own if k not in sub]
    extra = [k for k in sub if k not in own]
    assert not missing and not extra, f"state_dict mismatch: missing={missing[:5]} extra={extra[:5]}"
    module.load_state_dict(sub, strict=True)
    return module
