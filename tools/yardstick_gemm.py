#!/usr/bin/env python3
"""Yardstick only (NOT the product path): the 1x1-conv GEMM shapes of ResNet-50 (pixels x C_out x C_in, batch 256) through
zb_gemm (this library's tcgen05 kernel) and, for comparison, through torch.matmul with TF32 enabled (cuBLAS), both timed with CUDA
events around single launches with a 256 MB L2 flush in between.  Tells how far each shape is from what the hardware library
reaches on the same operands.  Usage: python tools/yardstick_gemm.py [--no-cublas]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenu_b200 import ZB_MATH_TF32, ops  # noqa: E402

SHAPES = [(802816, 256, 64), (802816, 64, 256), (802816, 128, 256), (200704, 512, 128), (200704, 128, 512), (200704, 256, 512),
          (50176, 1024, 256), (50176, 256, 1024), (50176, 512, 1024), (12544, 2048, 512), (12544, 512, 2048), (12544, 2048, 1024)]


def timeit(fn, flush, iters=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    use_cublas = "--no-cublas" not in sys.argv
    torch.backends.cuda.matmul.allow_tf32 = True
    ctx = ops.Context(math=ZB_MATH_TF32)
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    for (m, n, k) in SHAPES:
        a = torch.randn(m, k, device="cuda")
        w = torch.randn(n, k, device="cuda")      # [n][k]: B K-major (conv filter layout)
        wt = w.t().contiguous()                   # [k][n]: B MN-major (pointwise dgrad)
        c = torch.empty(m, n, device="cuda")
        r = {"m": m, "n": n, "k": k, "gflop": 2e-9 * m * n * k, "mbytes": 4e-6 * (m * k + n * k + m * n)}
        r["zb_nt_ms"] = timeit(lambda: ops.gemm(ctx, a, w, trans_b=True, c=c, math=ZB_MATH_TF32), flush)
        r["zb_nn_ms"] = timeit(lambda: ops.gemm(ctx, a, wt, c=c, math=ZB_MATH_TF32), flush)
        if use_cublas:
            r["cublas_nt_ms"] = timeit(lambda: torch.matmul(a, w.t(), out=c), flush)
            r["cublas_nn_ms"] = timeit(lambda: torch.matmul(a, wt, out=c), flush)
        for key in list(r):
            if key.endswith("_ms"):
                r[key.replace("_ms", "_tflops")] = round(r["gflop"] / r[key], 1)
                r[key] = round(r[key], 4)
        r["hbm_ms"] = round(r["mbytes"] / 6551.4, 4)
        print(json.dumps(r), flush=True)
    ctx.check()


if __name__ == "__main__":
    main()
