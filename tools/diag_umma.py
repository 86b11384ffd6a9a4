#!/usr/bin/env python3
"""Bring-up diagnostics for the tcgen05 path: runs every operand mode of the kernel on small problems and
prints error statistics against float64 numpy, one JSON line per case (never raises).  GPU box only.
Usage: python tools/diag_umma.py [out.json]"""
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import zenu_oracle as zo  # noqa: E402  (diagnostic tool, not product)
from zenu_b200 import ZB_MATH_TF32, ZB_NHWC, ops  # noqa: E402

RESULTS = []


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def nhwc(a):
    return np.ascontiguousarray(np.transpose(a, (0, 2, 3, 1)))


def nchw(a):
    return np.ascontiguousarray(np.transpose(a, (0, 3, 1, 2)))


def stats(name, got, ref, t):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    diff = np.abs(got - ref)
    r = {"case": name, "rel_l2": float(np.linalg.norm(diff) / (np.linalg.norm(ref) + 1e-300)),
         "max_abs": float(diff.max()), "ref_max": float(np.abs(ref).max()), "nan": int(np.isnan(got).sum()),
         "frac_bad": float((diff > 1e-2 * np.abs(ref).max()).mean()), "ms": round(t * 1e3, 3)}
    RESULTS.append(r)
    print(json.dumps(r), flush=True)


def run(name, fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        r = {"case": name, "error": repr(e)[:300]}
        RESULTS.append(r)
        print(json.dumps(r), flush=True)
        traceback.print_exc()


def main():
    ctx = ops.Context(math=ZB_MATH_TF32)
    rng = np.random.default_rng(0)

    def gemm_case(m, n, k, ta, tb):
        def f():
            a = rng.standard_normal((k, m) if ta else (m, k)).astype(np.float32)
            b = rng.standard_normal((n, k) if tb else (k, n)).astype(np.float32)
            ref = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
            A, B = dev(a), dev(b)
            torch.cuda.synchronize(); t0 = time.time()
            c = ops.gemm(ctx, A, B, ta, tb)
            ctx.check(); t = time.time() - t0
            stats(f"gemm m{m} n{n} k{k} ta{int(ta)} tb{int(tb)}", c.cpu().numpy(), ref, t)
        run(f"gemm m{m} n{n} k{k} ta{int(ta)} tb{int(tb)}", f)

    gemm_case(128, 64, 32, False, True)      # one tile, one k-block, K-major/K-major
    gemm_case(128, 64, 64, False, True)
    gemm_case(256, 128, 256, False, True)
    gemm_case(128, 64, 32, True, True)       # A MN-major
    gemm_case(128, 64, 32, False, False)     # B MN-major
    gemm_case(128, 64, 32, True, False)
    gemm_case(300, 200, 100, False, True)    # ragged
    gemm_case(300, 200, 100, True, False)
    gemm_case(512, 1000, 2048, False, True)  # fc-like, BN=256 tiles
    gemm_case(64, 512, 8192, False, True)    # split-K
    gemm_case(4096, 4096, 4096, False, True)

    def conv_case(tag, n, c, h, w, k, r, s, pad, stride, dil, which=("f", "d", "w")):
        x = rng.standard_normal((n, c, h, w)).astype(np.float32)
        wt = (rng.standard_normal((k, c, r, s)) / np.sqrt(c * r * s)).astype(np.float32)
        y_ref = zo.conv2d_fwd(x.astype(np.float64), wt.astype(np.float64), pad, stride, dil)
        dy = rng.standard_normal(y_ref.shape).astype(np.float32)
        X, W, DY = dev(nhwc(x)), dev(nhwc(wt)), dev(nhwc(dy))
        if "f" in which:
            def f():
                torch.cuda.synchronize(); t0 = time.time()
                y = ops.conv_fwd(ctx, X, W, pad, stride, dil, layout=ZB_NHWC)
                ctx.check(); t = time.time() - t0
                stats(tag + " fprop", nchw(y.cpu().numpy()), y_ref, t)
            run(tag + " fprop", f)
        if "d" in which:
            def f():
                ref = zo.conv2d_bkwd_data(dy.astype(np.float64), wt.astype(np.float64), x.shape, pad, stride, dil)
                torch.cuda.synchronize(); t0 = time.time()
                dx = ops.conv_bkwd_data(ctx, DY, W, X.shape, pad, stride, dil, layout=ZB_NHWC)
                ctx.check(); t = time.time() - t0
                stats(tag + " dgrad", nchw(dx.cpu().numpy()), ref, t)
            run(tag + " dgrad", f)
        if "w" in which:
            def f():
                ref = zo.conv2d_bkwd_filter(dy.astype(np.float64), x.astype(np.float64), wt.shape, pad, stride, dil)
                torch.cuda.synchronize(); t0 = time.time()
                dw = ops.conv_bkwd_weight(ctx, DY, X, W.shape, pad, stride, dil, layout=ZB_NHWC)
                ctx.check(); t = time.time() - t0
                stats(tag + " wgrad", nchw(dw.cpu().numpy()), ref, t)
            run(tag + " wgrad", f)

    conv_case("1x1 c64 k64 8x8", 2, 64, 8, 8, 64, 1, 1, 0, 1, 1)
    conv_case("3x3 c32 k32 8x8 n2", 2, 32, 8, 8, 32, 3, 3, 1, 1, 1)
    conv_case("3x3 c64 k128 14x14", 2, 64, 14, 14, 128, 3, 3, 1, 1, 1)
    conv_case("3x3 s2 c64 k64 15x17", 2, 64, 15, 17, 64, 3, 3, 1, 2, 1)
    conv_case("1x1 s2 c64 k96 16x16", 2, 64, 16, 16, 96, 1, 1, 0, 2, 1)
    conv_case("3x3 d2 c32 k40 12x12", 1, 32, 12, 12, 40, 3, 3, 2, 1, 2)
    conv_case("5x5 c32 k32 8x8", 2, 32, 8, 8, 32, 5, 5, 2, 1, 1)
    conv_case("3x3 c64 k64 56x56 n4", 4, 64, 56, 56, 64, 3, 3, 1, 1, 1)
    conv_case("1x1 c256 k64 56x56 n4", 4, 256, 56, 56, 64, 1, 1, 0, 1, 1)
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/diag_umma.json"
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out, "w") as f:
        json.dump(RESULTS, f, indent=1)


if __name__ == "__main__":
    main()
