#!/usr/bin/env python3
"""BASELINE.json configs[4]: BatchNorm2d (+ReLU, +residual add) forward / backward bandwidth sweep, NHWC f32, tensors from 1 MB to
2 GB, against the measured HBM copy bandwidth.  Algorithmic bytes / element: fwd 12 (+4 with residual); bwd 20 (BN+ReLU, mask
recomputed from x); with residual 24.25 through the 1-bit ReLU mask of the fused forward (what the model runs), 28 through y.
Usage: python tools/bench_bn.py [--out gpurun_out/bn_sweep.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenu_b200 import ZB_NHWC, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/bn_sweep.json")
    a = ap.parse_args()
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            hbm = float(json.load(f)["hbm_gbs"])
        src = "measured"
    except Exception:  # noqa: BLE001
        hbm, src = 6650.0, "fallback"
    ctx = ops.Context()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    rows = []
    for c in (64, 256, 1024):
        for mb in (1, 4, 16, 64, 256, 822, 2048):
            elems = mb * 1024 * 1024 // 4
            hw = 28 if mb >= 16 else 7
            n = max(1, elems // (c * hw * hw))
            shape = (n, hw, hw, c)
            x = torch.randn(shape, device="cuda")
            res = torch.randn(shape, device="cuda")
            dy = torch.randn(shape, device="cuda")
            sc, bi = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
            rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
            numel = x.numel()
            out = {}

            def timeit(fn):
                for _ in range(2):
                    fn()
                ts = []
                for _ in range(a.iters):
                    flush.fill_(1.0)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); fn(); e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                return sorted(ts)[len(ts) // 2]
            y, sm, si = ops.batch_norm_2d_forward_train(ctx, 0.9, x, sc, bi, rm, rv, layout=ZB_NHWC, relu=True)
            y2, sm2, si2 = ops.batch_norm_2d_forward_train(ctx, 0.9, x, sc, bi, rm, rv, layout=ZB_NHWC, residual=res, relu=True)
            y3, sm3, si3, mask3 = ops.batch_norm_2d_forward_train_masked(ctx, 0.9, x, sc, bi, rm, rv, residual=res)
            cases = {
                "fwd+relu": (12, lambda: ops.batch_norm_2d_forward_train(ctx, 0.9, x, sc, bi, rm, rv, layout=ZB_NHWC, relu=True)),
                "fwd+relu+res": (16, lambda: ops.batch_norm_2d_forward_train(ctx, 0.9, x, sc, bi, rm, rv, layout=ZB_NHWC, residual=res, relu=True)),
                "bwd+relu": (20, lambda: ops.batch_norm_2d_relu_backward(ctx, x, dy, sc, bi, sm, si, layout=ZB_NHWC)),
                # what the model runs: the 1-bit ReLU mask written by the fused forward (x, dy, mask read; masked gradient written
                # and re-read; x re-read; dx written: 6 passes + 2/32)
                "bwd+relu+res": (24.25, lambda: ops.batch_norm_2d_backward_masked(ctx, x, dy, sc, sm3, si3, mask3)),
                # the y-based form of the reference's separate nodes (7 passes)
                "bwd+relu+res (y)": (28, lambda: ops.batch_norm_2d_backward(ctx, x, dy, sc, sm2, si2, layout=ZB_NHWC, y=y2, want_residual_grad=True)),
                "fwd+relu+res+mask": (16.125, lambda: ops.batch_norm_2d_forward_train_masked(ctx, 0.9, x, sc, bi, rm, rv, residual=res)),
            }
            for name, (bpe, fn) in cases.items():
                ms = timeit(fn)
                gbs = numel * bpe / ms / 1e6
                out[name] = {"ms": ms, "gbs": gbs, "frac_of_hbm_peak": gbs / hbm}
            ctx.check()
            rows.append({"c": c, "shape": list(shape), "mbytes": numel * 4 / 1e6, **out})
            print(f"C={c:<5d} {numel * 4 / 1e6:8.1f} MB |" + "".join(f" {k} {v['gbs']:6.0f} GB/s {v['frac_of_hbm_peak'] * 100:5.1f}% |" for k, v in out.items()), flush=True)
            del x, res, dy, y, y2, y3
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump({"hbm_gbs": hbm, "peak_source": src, "l2": "256 MB scratch write between launches", "rows": rows}, f, indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
