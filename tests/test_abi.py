"""CPU: the C-ABI library loads without a GPU and exports exactly the symbols include/mqdet_b200.h declares."""
import os
import re

from util import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "mqdet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mqdet_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    from mqdet_b200 import _lib
    lib = _lib.load()  # binds every name in SIGNATURES; AttributeError if one is missing
    declared = _declared()
    assert declared, "no declarations parsed"
    assert sorted(_lib.SIGNATURES) == declared, (sorted(set(declared) ^ set(_lib.SIGNATURES)))
    for name in declared:
        assert hasattr(lib, name)
    assert lib.mqdet_version() >= 100
    assert lib.mqdet_ml_nms_workspace_bytes(5000) > 5000 * 79 * 8


def test_gemm_args_struct_matches_header_field_order():
    from mqdet_b200 import _lib
    src = open(os.path.join(ROOT, "include", "mqdet_b200.h")).read()
    body = src[src.index("typedef struct mqdet_gemm_args {"):src.index("} mqdet_gemm_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        parts = decl.replace("*", " ").split(",")
        names.append(parts[0].split()[-1])
        names.extend(p.strip() for p in parts[1:])
    assert names == [f[0] for f in _lib.GemmArgs._fields_]


def test_product_path_never_imports_the_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "mqdet_b200")):
        for f in files:
            if f.endswith(".py") and re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(d, f)).read(), re.M):
                bad.append(f)
    assert not bad, bad


def test_no_cpu_fallback():
    import pytest
    import torch
    from mqdet_b200 import ops
    from mqdet_b200._lib import MqdetError
    with pytest.raises(MqdetError):
        ops.layernorm(torch.zeros(2, 8), torch.ones(8), torch.zeros(8))


def test_argument_validation_errors_without_gpu():
    """Every entry point validates its arguments BEFORE touching CUDA: a bad call returns a negative code and a message in
    mqdet_last_error() (mirrors the reference's AT_ERROR / TORCH_CHECK behaviour), also on a machine without a GPU."""
    import ctypes
    from mqdet_b200 import _lib
    lib = _lib.load()
    nul = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)  # never dereferenced: the checks fail first

    def expect(rc, needle):
        assert rc < 0
        msg = lib.mqdet_last_error().decode()
        assert needle in msg, msg

    g = _lib.GemmArgs()
    expect(lib.mqdet_gemm_f16(ctypes.byref(g), 0, nul), "null pointer")
    g.A = g.B = g.C = 16
    g.M, g.N, g.K = 4, 4, 0
    expect(lib.mqdet_gemm_f16(ctypes.byref(g), 0, nul), "empty problem")
    g.K, g.lda, g.ldb, g.ldc, g.nb1, g.nb2 = 12, 12, 12, 4, 1, 1   # K % 8 != 0
    expect(lib.mqdet_gemm_f16(ctypes.byref(g), 0, nul), "multiples of 8")
    g.K = g.lda = g.ldb = 16
    expect(lib.mqdet_gemm_f16(ctypes.byref(g), 7, nul), "unknown impl")
    expect(lib.mqdet_layernorm(nul, 0, 8, nul, nul, 1e-5, 1, 8, nul, nul, 8, 0, nul), "null pointer")
    expect(lib.mqdet_colsoftmax_transposed(one, 1, 10, 12, one, 16, one, nul), "T%8")          # T % 8 != 0
    expect(lib.mqdet_colstats_rowsoftmax(one, 1, 10, 128, nul, 1, 0.0, 0.0, one, nul), "T must be 256")
    expect(lib.mqdet_swin_window_attn(one, one, one, 1, 14, 14, 3, 8, 0, 1.0, one, nul), "window 7 (Swin-T) or 12 (Swin-L)")
    expect(lib.mqdet_biattn_image(one, 8, 8, one, 8, 8, 8, nul, 0, one, 8, 8, 8, nul, nul, nul, 0, 0, nul, 0.0, one, 8, 8, one,
                                  one, 1, 8, 100, 100, nul), "T % 8 == 0")
    expect(lib.mqdet_biattn_text_vn(one, 8, 8, 8, one, 8, 8, 8, one, 8, 0, 8, nul, nul, 0, 0.0, one, 8, 8, 8, 1, 1, 256, 100, nul),
           "null pointer")
    expect(lib.mqdet_gather_detections(one, one, one, one, one, 1, 256, 128, 64, one, nul), "det_rows")
    expect(lib.mqdet_dcn_cols(one, nul, 0, one, 5, 1, 128, 1, one, nul), "C must be 256")
    expect(lib.mqdet_dcn_conv(one, nul, 0, one, 5, 1, 128, 1, one, one, one, one, nul), "C must be 256")
    expect(lib.mqdet_dcn_conv(one, nul, 0, one, 5, 1, 256, 4, one, one, one, one, nul), "1..3 jobs")
    expect(lib.mqdet_conv3x3_small(one, one, one, one, 5, 1, 256, 40, one, 64, nul), "O <= 32")
    expect(lib.mqdet_biattn_text(one, 8, 8, 8, one, 8, 8, 8, one, 8, 8, 8, one, 0.0, one, 8, 8, 8, 1, 1, 256, 100, 104, 128,
                                 nul), "head dim must be 256")
    expect(lib.mqdet_contrastive_mask(one, one, 1, 1, 300, 256, nul), "bad args")                # Tmax < T
