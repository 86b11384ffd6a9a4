#!/usr/bin/env python3
"""Diagnostic twin of tests/test_gpu_model.py: prints loss / per-parameter gradient errors instead of asserting.
GPU box only (diagnostic tool: uses the oracle as the checker)."""
import json
import os
import sys
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zenu_b200 as pkg  # noqa: E402
from oracle import zenu_oracle_model as zm  # noqa: E402
from tests.test_gpu_model import CASES, batch, load_params, rel  # noqa: E402
from zenu_b200 import nn, ops  # noqa: E402


def run(arch, n, hw, classes, math, opt, ltol, gtol, fused):
    ctx = ops.Context(math=pkg.ZB_MATH_TF32 if math == "tf32" else pkg.ZB_MATH_FP32)
    params = zm.init_params(arch, classes, seed=42)
    oracle = zm.OracleModel(arch, classes, {k: v.copy() for k, v in params.items()})
    model = nn.Model(ctx, arch, classes, fused=fused, seed=1)
    load_params(model, params)
    kw = dict(kind=opt, lr=0.01 if opt == "sgd" else 1e-3, weight_decay=0.01 if opt == "adamw" else 0.0)
    model.set_optimizer(**kw)
    x, t = batch(n, hw, classes, 1234)
    X, T = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
    loss_ref, grads_ref = oracle.forward_backward(x, t)
    loss = model.forward_backward(X, T)
    ctx.check()
    named = model.named_parameters()
    errs = []
    for name, g_ref in grads_ref.items():
        g = named[name]["grad"]
        if name.endswith("conv2d.filter"):
            g = model.filter_to_kcrs(g)
        errs.append((rel(g.cpu().numpy(), g_ref), name, float(np.abs(g_ref).max())))
    errs.sort(reverse=True)
    out = {"case": [arch, n, hw, classes, math, opt, fused], "loss": float(loss.item()), "loss_ref": float(loss_ref),
           "ltol": ltol, "gtol": gtol, "worst": [(round(e, 6), nm, mx) for e, nm, mx in errs[:6]]}
    oracle.update(grads_ref, **kw)
    model.update()
    curve = []
    for _ in range(2):
        curve.append((model.train_step(X, T, read_loss=True), oracle.train_step(x, t, **kw)))
    out["curve"] = curve
    print(json.dumps(out), flush=True)
    model.close()
    ctx.close()


for case in CASES:
    for fused in (True, False):
        if case[0] == "resnet50" and not fused:
            continue
        try:
            run(*case, fused)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"case": list(case) + [fused], "error": repr(e)[:400]}), flush=True)
            traceback.print_exc()
