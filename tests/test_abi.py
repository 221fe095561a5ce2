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
