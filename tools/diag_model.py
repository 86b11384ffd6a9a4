#!/usr/bin/env python3
"""Diagnostic twin of tests/test_gpu_model.py: prints loss / per-parameter gradient errors instead of asserting.
GPU box only (diagnostic tool: uses the oracle as the checker)."""
import json
import os
import os as _os
import sys
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zenu_b200 as pkg  # noqa: E402
from oracle import zenu_oracle_model as zm  # noqa: E402
from tests.test_gpu_model import CASES, batch, load_params, rel  # noqa: E402
from zenu_b200 import nn, ops  # noqa: E402


def run(arch, n, hw, classes, math, opt, ltol, gtol, fused, rnd=None):
    ctx = ops.Context(math=pkg.ZB_MATH_TF32 if math == "tf32" else pkg.ZB_MATH_FP32)
    params = zm.init_params(arch, classes, seed=42)
    oracle = zm.OracleModel(arch, classes, {k: v.copy() for k, v in params.items()}, operand_round=rnd)
    model = nn.Model(ctx, arch, classes, fused=fused, seed=1)
    load_params(model, params)
    kw = dict(kind=opt, lr=float(_os.environ.get("DIAG_LR", "0.01")) if opt == "sgd" else 1e-3, weight_decay=0.01 if opt == "adamw" else 0.0)
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
    fwd = []
    for name, ent in named.items():
        if name.endswith("batch_norm_2d.variance") or name.endswith("batch_norm_2d.mean"):
            fwd.append((name.replace(".batch_norm_2d", ""), round(rel(ent["data"].cpu().numpy(), oracle.p[name]), 7)))
    gall = np.concatenate([(model.filter_to_kcrs(named[k]["grad"]) if k.endswith("conv2d.filter") else named[k]["grad"]).cpu().numpy().ravel()
                           for k in grads_ref])
    rall = np.concatenate([grads_ref[k].ravel() for k in grads_ref])
    whole = rel(gall, rall)
    cos = float(np.dot(gall.astype(np.float64), rall.astype(np.float64)) / (np.linalg.norm(gall.astype(np.float64)) * np.linalg.norm(rall.astype(np.float64))))
    tail = [(round(e, 6), nm) for e, nm, _ in errs if nm.startswith("fc.") or nm.startswith("linear2.") or nm.startswith("layer4.") and "conv2d" in nm][:6]
    errs.sort(reverse=True)
    out = {"case": [arch, n, hw, classes, math, opt, fused, rnd], "loss": float(loss.item()), "loss_ref": float(loss_ref),
           "ltol": ltol, "gtol": gtol, "fwd_stats": fwd[-2:], "whole_grad_rel": whole, "cos": cos, "tail": tail, "worst": [(round(e, 6), nm, mx) for e, nm, mx in errs[:6]]}
    oracle.update(grads_ref, **kw)
    model.update()
    curve = []
    for _ in range(2):
        curve.append((model.train_step(X, T, read_loss=True), oracle.train_step(x, t, **kw)))
    out["curve"] = curve
    print(json.dumps(out), flush=True)
    model.close()
    ctx.close()


ONLY = _os.environ.get("DIAG_ONLY")
if _os.environ.get("DIAG_R50"):
    CASES = [("resnet50", int(b), int(h), 8, "tf32", "sgd", 5e-3, 5e-2) for b, h in (x.split("x") for x in _os.environ["DIAG_R50"].split(","))]
for case in CASES:
    if ONLY and case[0] != ONLY:
        continue
    for fused, rnd in ((True, None), (True, "rne"), (False, None)):
        if (case[0] == "resnet50" and not fused) or (rnd and case[4] != "tf32"):
            continue
        try:
            run(*case, fused, rnd)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"case": list(case) + [fused], "error": repr(e)[:400]}), flush=True)
            traceback.print_exc()
