#!/usr/bin/env python3
"""BASELINE.json configs[3]: Conv2d fprop / dgrad / wgrad microbench over the 23 distinct ResNet-50 layer shapes (N = 256, f32
storage, TF32 tensor-core math) against the per-layer roofline min(TF32 peak, AI x HBM bandwidth).
Timing: CUDA events around `iters` back-to-back launches after warm-up; a 256 MB scratch write between launches flushes L2.
--layout nchw times the reference-contract path (NCHW activations, KCRS filters: what a port holding zenu-matrix `Matrix` tensors
passes, zenu-matrix/src/nn/conv/interface.rs:270-281) instead of the backend's native NHWC / KRSC.
Usage: python tools/bench_conv.py [--batch 256] [--iters 5] [--only 3x3] [--layout nhwc|nchw] [--out gpurun_out/conv_sweep.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenu_b200 import ZB_MATH_TF32, ZB_NCHW, ZB_NHWC, ops  # noqa: E402

# C_in, H_in, C_out, k, stride, pad, count  (SURVEY Table 8d-1)
SHAPES = [(3, 224, 64, 7, 2, 3, 1), (64, 56, 64, 1, 1, 0, 1), (64, 56, 64, 3, 1, 1, 3), (64, 56, 256, 1, 1, 0, 4), (256, 56, 64, 1, 1, 0, 2),
          (256, 56, 128, 1, 1, 0, 1), (128, 56, 128, 3, 2, 1, 1), (128, 28, 512, 1, 1, 0, 4), (256, 56, 512, 1, 2, 0, 1), (512, 28, 128, 1, 1, 0, 3),
          (128, 28, 128, 3, 1, 1, 3), (512, 28, 256, 1, 1, 0, 1), (256, 28, 256, 3, 2, 1, 1), (256, 14, 1024, 1, 1, 0, 6), (512, 28, 1024, 1, 2, 0, 1),
          (1024, 14, 256, 1, 1, 0, 5), (256, 14, 256, 3, 1, 1, 5), (1024, 14, 512, 1, 1, 0, 1), (512, 14, 512, 3, 2, 1, 1), (512, 7, 2048, 1, 1, 0, 3),
          (1024, 14, 2048, 1, 2, 0, 1), (2048, 7, 512, 1, 1, 0, 2), (512, 7, 512, 3, 1, 1, 2)]


def peaks():
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]) / 2.0, "measured (bf16 burst / 2)"
    except Exception:  # noqa: BLE001
        return 6650.0, 1590.0 / 2.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--layout", default="nhwc", choices=["nhwc", "nchw"])
    ap.add_argument("--out", default="gpurun_out/conv_sweep.json")
    a = ap.parse_args()
    hbm, tf32, src = peaks()
    ctx = ops.Context(math=ZB_MATH_TF32)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    rows = []
    tot = {"fprop": [0.0, 0.0], "dgrad": [0.0, 0.0], "wgrad": [0.0, 0.0]}
    for (c, h, k, r, s, p, cnt) in SHAPES:
        tag = f"{r}x{r}"
        if a.only and a.only != tag and a.only != f"c{c}":
            continue
        n = a.batch
        ho = (h + 2 * p - r) // s + 1
        L = ZB_NHWC if a.layout == "nhwc" else ZB_NCHW
        if a.layout == "nhwc":
            x = torch.randn((n, h, h, c), device="cuda")
            w = torch.randn((k, r, r, c), device="cuda") * (2.0 / (c * r * r)) ** 0.5
            dy = torch.randn((n, ho, ho, k), device="cuda")
        else:
            x = torch.randn((n, c, h, h), device="cuda")
            w = torch.randn((k, c, r, r), device="cuda") * (2.0 / (c * r * r)) ** 0.5
            dy = torch.randn((n, k, ho, ho), device="cuda")
        flops = 2.0 * n * ho * ho * k * c * r * r
        # algorithmic bytes: a strided pointwise conv touches only every stride-th pixel of x in fprop / wgrad (the rest of x is never
        # needed); its dgrad still has to WRITE all of dx (zeros where no output pixel reaches)
        x_touched = n * h * h * c / (s * s) if (r == 1 and s > 1) else n * h * h * c
        byts_by = {"fprop": 4.0 * (x_touched + k * c * r * r + n * ho * ho * k), "wgrad": 4.0 * (x_touched + k * c * r * r + n * ho * ho * k),
                   "dgrad": 4.0 * (n * h * h * c + k * c * r * r + n * ho * ho * k)}
        byts = byts_by["dgrad"]
        ideal_by = {nm: max(flops / (tf32 * 1e12), b / (hbm * 1e9)) * 1e3 for nm, b in byts_by.items()}
        ideal_ms = ideal_by["fprop"]
        rec = {"c": c, "hw": h, "k": k, "r": r, "stride": s, "count": cnt, "gflop": flops / 1e9, "mbytes": byts_by["fprop"] / 1e6,
               "mbytes_dgrad": byts / 1e6, "bound": "tensor" if flops / (tf32 * 1e12) > byts_by["fprop"] / (hbm * 1e9) else "hbm",
               "ideal_ms": ideal_ms, "ideal_ms_dgrad": ideal_by["dgrad"], "layout": a.layout}
        fns = {"fprop": lambda: ops.conv_fwd(ctx, x, w, p, s, 1, layout=L),
               "dgrad": lambda: ops.conv_bkwd_data(ctx, dy, w, x.shape, p, s, 1, layout=L),
               "wgrad": lambda: ops.conv_bkwd_weight(ctx, dy, x, w.shape, p, s, 1, layout=L)}
        for name, fn in fns.items():
            for _ in range(2):
                fn()
            ts = []
            for _ in range(a.iters):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            rec[name] = {"ms": ms, "tflops": flops / ms / 1e9, "gbs": byts_by[name] / ms / 1e6, "frac_of_tf32_peak": flops / ms / 1e9 / tf32,
                         "frac_of_layer_roofline": ideal_by[name] / ms}
            tot[name][0] += ms * cnt
            tot[name][1] += ideal_by[name] * cnt
        ctx.check()
        rows.append(rec)
        print(f"c{c:<5d}hw{h:<4d}k{k:<5d}{r}x{r} s{s} x{cnt}  {rec['bound']:6s} ideal {ideal_ms:6.3f} ms |" +
              "".join(f" {nm} {rec[nm]['ms']:6.3f} ms {rec[nm]['tflops']:6.1f} TF/s {rec[nm]['frac_of_layer_roofline'] * 100:5.1f}% |" for nm in fns), flush=True)
        del x, w, dy
    summary = {nm: {"ms_per_net": v[0], "roofline_ms": v[1], "frac": v[1] / v[0] if v[0] else None} for nm, v in tot.items()}
    print(json.dumps(summary))
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump({"peaks": {"hbm_gbs": hbm, "tf32_tflops": tf32, "source": src}, "batch": a.batch, "layout": a.layout, "l2": "256 MB scratch write between launches",
                   "layers": rows, "per_network": summary}, f, indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
