#!/usr/bin/env python3
"""Per-layer check of the tcgen05 conv path against the oracle fed with tf32-rounded operands (zo.tf32_round): with the
operand rounding modelled, what is left is fp32 accumulation order, so errors far above ~1e-5 point at a kernel bug
rather than at TF32.  GPU box only (diagnostic: uses the oracle as the checker).
Usage: python tools/diag_conv_layers.py [arch] [batch] [hw]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import zenu_oracle as zo  # noqa: E402
from oracle import zenu_oracle_model as zm  # noqa: E402
from zenu_b200 import ZB_MATH_TF32, ZB_NHWC, ops  # noqa: E402


def nhwc(a):
    return np.ascontiguousarray(np.transpose(a, (0, 2, 3, 1)))


def nchw(a):
    return np.ascontiguousarray(np.transpose(a, (0, 3, 1, 2)))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm((a - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30))


def layers(arch, hw):
    if arch == "small_cnn":
        return [("conv1", 3, 32, 3, 1, 1, hw), ("conv2", 32, 64, 3, 1, 1, hw)]
    blocks, _ = zm._resnet_plan(18 if arch == "resnet18" else 50)
    out = [("conv1", 3, 64, 7, 2, 3, hw)]
    h = (hw + 6 - 7) // 2 + 1
    h = (h + 2 - 3) // 2 + 1
    for name, convs, down in blocks:
        hin = h
        for i, (ci, co, k, stride, pad) in enumerate(convs):
            out.append((f"{name}.conv{i + 1}", ci, co, k, stride, pad, h))
            h = (h + 2 * pad - k) // stride + 1
        if down:
            out.append((f"{name}.down", down[0], down[1], 1, down[3], 0, hin))
    return out


def main():
    arch = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    hw = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    mode = sys.argv[4] if len(sys.argv) > 4 else "rne"
    ctx = ops.Context(math=ZB_MATH_TF32)
    zo.use_openblas()
    rng = np.random.default_rng(0)
    seen = set()
    for name, ci, co, k, stride, pad, h in layers(arch, hw):
        key = (ci, co, k, stride, pad, h)
        if key in seen:
            continue
        seen.add(key)
        x = rng.standard_normal((n, ci, h, h)).astype(np.float32)
        w = (rng.standard_normal((co, ci, k, k)) * np.sqrt(2.0 / (ci * k * k))).astype(np.float32)
        xr, wr = zo.tf32_round(x, mode), zo.tf32_round(w, mode)
        y_ref = zo.conv2d_fwd(xr, wr, pad, stride, 1)
        dy = rng.standard_normal(y_ref.shape).astype(np.float32)
        dyr = zo.tf32_round(dy, mode)
        dx_ref = zo.conv2d_bkwd_data(dyr, wr, x.shape, pad, stride, 1)
        dw_ref = zo.conv2d_bkwd_filter(dyr, xr, w.shape, pad, stride, 1)
        X, W, DY = dev(nhwc(x)), dev(nhwc(w)), dev(nhwc(dy))
        y = ops.conv_fwd(ctx, X, W, pad, stride, 1, layout=ZB_NHWC)
        dx = ops.conv_bkwd_data(ctx, DY, W, X.shape, pad, stride, 1, layout=ZB_NHWC)
        dw = ops.conv_bkwd_weight(ctx, DY, X, W.shape, pad, stride, 1, layout=ZB_NHWC)
        ctx.check()
        r = {"layer": name, "shape": key, "fprop": rel(nchw(y.cpu().numpy()), y_ref), "dgrad": rel(nchw(dx.cpu().numpy()), dx_ref),
             "wgrad": rel(nchw(dw.cpu().numpy()), dw_ref)}
        r["flag"] = "BAD" if max(r["fprop"], r["dgrad"], r["wgrad"]) > 2e-5 else "ok"
        print(json.dumps(r), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
