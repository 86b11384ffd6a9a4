#!/usr/bin/env python3
"""Target for `ncu --profile-from-start off`: one launch each of three 1x1-conv GEMM shapes (pixels x C_out x C_in) of the 14x14 / 7x7
stages through zb_gemm, inside cudaProfilerStart/Stop.  Usage: ncu ... python tools/ncu_gemm.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenu_b200 import ZB_MATH_TF32, ops  # noqa: E402

ctx = ops.Context(math=ZB_MATH_TF32)
probs = []
for (m, n, k) in [(50176, 256, 1024), (50176, 1024, 256), (12544, 512, 2048)]:
    probs.append((torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda"), torch.empty(m, n, device="cuda")))
for a, w, c in probs:
    ops.gemm(ctx, a, w, trans_b=True, c=c, math=ZB_MATH_TF32)
flush = torch.empty(64 * 1024 * 1024, device="cuda")
torch.cuda.synchronize()
for a, w, c in probs:
    flush.fill_(1.0)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ops.gemm(ctx, a, w, trans_b=True, c=c, math=ZB_MATH_TF32)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
ctx.check()
