"""First-contact diagnostics on the B200 box: prints error magnitudes instead of asserting, so one gpurun call tells
which layer of the stack (TMA map, tcgen05 descriptors, epilogue, GCP kernels) is wrong."""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from mqdet_b200 import ops

res = {}
dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))


def run(name, fn):
    t = time.time()
    try:
        r = fn()
        torch.cuda.synchronize()
        res[name] = r
        print(f"[{name}] {r} ({time.time()-t:.2f}s)", flush=True)
    except Exception as e:
        res[name] = "EXC: " + repr(e)
        print(f"[{name}] EXC {e!r}", flush=True)
        traceback.print_exc()
        try:
            torch.cuda.synchronize()
        except Exception as e2:
            print("device is wedged:", e2)
            json.dump(res, open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w"), indent=1)
            sys.exit(3)


def gemm_case(M, N, K, impl):
    g = torch.Generator().manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g) * 0.5).half().to(dev)
    b = (torch.randn(N, K, generator=g) * 0.5).half().to(dev)
    out = ops.gemm(a, b, out_dtype=torch.float32, impl=impl)
    ref = a.float() @ b.float().T
    err = (out - ref).abs().max().item()
    # row/col structure of the error helps telling descriptor bugs from epilogue bugs
    bad = ((out - ref).abs() > 1e-2 * ref.abs().max()).float()
    return dict(err=err, ref=ref.abs().max().item(), bad_frac=bad.mean().item(),
                bad_rows=int((bad.sum(1) > 0).sum()), bad_cols=int((bad.sum(0) > 0).sum()))


os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
run("simt_128", lambda: gemm_case(128, 128, 64, ops.IMPL_SIMT))
for (M, N, K) in [(128, 32, 64), (128, 64, 64), (128, 128, 64), (128, 256, 64), (128, 128, 128), (128, 128, 256),
                  (256, 256, 768), (2048, 512, 768), (2048, 3072, 768), (300, 200, 136), (22400, 256, 256)]:
    run(f"tc_{M}x{N}x{K}", lambda: gemm_case(M, N, K, ops.IMPL_TCGEN05))


def timing():
    out = {}
    for (M, N, K) in [(2048, 3072, 768), (2048, 768, 3072), (22400 * 8, 2048, 256), (22400 * 8, 256, 2048), (8192, 8192, 8192)]:
        a = torch.randn(M, K, device=dev).half()
        b = torch.randn(N, K, device=dev).half()
        c = torch.empty(M, N, device=dev, dtype=torch.float16)
        for _ in range(3):
            ops.gemm(a, b, out=c)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out[f"{M}x{N}x{K}"] = dict(ms=ms, tflops=2 * M * N * K / ms / 1e9)
        for _ in range(3):
            torch.matmul(a, b.T, out=c)
        e0.record()
        for _ in range(10):
            torch.matmul(a, b.T, out=c)
        e1.record()
        torch.cuda.synchronize()
        out[f"{M}x{N}x{K}"]["cublas_tflops"] = 2 * M * N * K / (e0.elapsed_time(e1) / 10) / 1e9
    return out


run("gemm_timing", timing)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w"), indent=1)
