"""fp16 operand cache for module parameters.

Parameters stay fp32 nn.Parameters under the reference's names (state_dict compatible); the tensor-core GEMMs read
fp16 copies that are (re)built lazily whenever the parameter's storage or version counter changes.
"""
import weakref

import torch

from .. import ops

_cache = {}


def w16(param, rows=None, view=None):
    """fp16, contiguous copy of ``param`` (optionally a row slice ``rows=(lo, hi)``, optionally reshaped to ``view``),
    cached per parameter version.  Always pass the nn.Parameter itself (not a temporary view) so the cache can hit."""
    if view is not None:
        return w16(param, rows).view(*view)
    key = (id(param), rows)
    ent = _cache.get(key)
    ver = (param.data_ptr(), param._version, param.device)
    if ent is not None and ent[0] == ver and ent[2]() is param:
        return ent[1]
    src = param.detach()
    if rows is not None:
        src = src[rows[0]:rows[1]]
    if src.dtype == torch.float16:
        h = src.contiguous()
    else:
        h = ops.cast_f16(src.float().contiguous())
    # weak reference: a recycled id()/data_ptr of a dead parameter can never alias a live one, and the fp16 copy is
    # dropped together with the parameter
    _cache[key] = (ver, h, weakref.ref(param, lambda _r, k=key: _cache.pop(k, None)))
    return h


def f32(param):
    """fp32 contiguous view of a parameter (LayerNorm affine, biases, gates)."""
    t = param.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


def clear():
    _cache.clear()


_cache_t = {}


def wT16(param):
    """fp16 TRANSPOSE of a 2-D parameter [out, in] -> [in, out rounded up to 8] (zero padded): the B operand of an activation
    gradient dX = dY W (K = out).  Cached per parameter version like ``w16`` (frozen weights are transposed once)."""
    key = id(param)
    ent = _cache_t.get(key)
    ver = (param.data_ptr(), param._version, param.device)
    if ent is not None and ent[0] == ver and ent[2]() is param:
        return ent[1]
    src = param.detach()
    h = ops.transpose_cast(src if src.dtype in (torch.float16, torch.float32) and src.is_contiguous() else src.float().contiguous())
    _cache_t[key] = (ver, h, weakref.ref(param, lambda _r, k=key: _cache_t.pop(k, None)))
    return h
