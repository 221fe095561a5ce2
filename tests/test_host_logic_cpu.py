"""CPU: the HOST-SIDE logic of the modules that were composed from existing kernels this round (GroundingDINO transformer / model /
extract_query, the K-split explicit BiAttention path, the training forwards / backwards) run over a torch test double of the C-ABI ops
(tests/hostlogic/ops_double.py) and compared with the oracle — each case in its own process, because the double monkey-patches
``mqdet_b200.ops`` and ``Tensor.is_cuda``.  This checks views / strides / masks / caches / operand order, NOT the kernels (``-m gpu``)."""
import os
import subprocess
import sys

import pytest

from util import ROOT

CASES = ["transformer", "ref_signatures", "biattn_split", "gcp_bwd", "bert_bwd", "preselect_bwd", "lang_train", "extract_query", "model"]


@pytest.mark.parametrize("case", CASES)
def test_host_logic(case):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), OMP_NUM_THREADS=os.environ.get("OMP_NUM_THREADS", "8"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "hostlogic", "run.py"), case], capture_output=True, text=True, timeout=1500,
                       env=env, cwd=ROOT)
    assert r.returncode == 0 and f"PASS {case}" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
