"""mqdet_b200 — B200-native (sm_100a) implementation of MQ-Det's multi-modal query forward pass.

Host code is Python/PyTorch plumbing over hand-written CUDA kernels behind a C ABI (include/mqdet_b200.h,
mqdet_b200/csrc).  The module tree mirrors the slice of ``maskrcnn_benchmark.modeling`` on the hot path
 with the reference's class names, constructor arguments, parameter names and forward signatures.
"""
from . import _lib, ops  # noqa: F401

__all__ = ["_lib", "ops"]
