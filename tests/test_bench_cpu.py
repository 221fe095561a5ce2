"""CPU: the `--impl reference` arm of bench.py (the CPU oracle on the host cores) — contract of the JSON line, and that ranks
other than 0 exit quietly under torchrun-style environments."""
import json
import os
import subprocess
import sys

from util import ROOT


def _run(env_extra, *args, timeout=600):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_json_line():
    r = _run({"RANK": "0", "WORLD_SIZE": "1", "MQDET_CPU_THREADS": "8"}, "--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    line = [x for x in r.stdout.splitlines() if x.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["metric"].startswith("images/sec MQ-GLIP-T") and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and abs(d["value"] * d["ms_per_step"] / 1e3 - 1.0) < 1e-6
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 8 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
