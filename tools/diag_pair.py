#!/usr/bin/env python3
"""Bring-up diagnostics for the CTA-pair (cluster of 2, TMA-multicast B) variant of the tcgen05 GEMM kernel: GEMMs in all four
transpose forms with and without ring wrap-around, error map per (128-row block, 32-column group) against float64.  GPU box only.
Usage: ZENU_B200_PAIR=1 python tools/diag_pair.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenu_b200 import ZB_MATH_TF32, ops  # noqa: E402


def main():
    ctx = ops.Context(math=ZB_MATH_TF32)
    g = torch.Generator().manual_seed(7)
    cases = []
    for (ta, tb) in [(False, True), (False, False), (True, False), (True, True)]:
        for (m, n, k) in [(256, 128, 128), (256, 128, 256), (243, 128, 256), (256, 256, 96), (256, 256, 512), (512, 256, 1024),
                          (1024, 512, 2048), (256 * 40, 256, 256)]:
            cases.append((m, n, k, ta, tb))
    for (m, n, k, ta, tb) in cases:
        a = torch.randn((k, m) if ta else (m, k), generator=g)
        b = torch.randn((n, k) if tb else (k, n), generator=g)
        ref = (a.double().T if ta else a.double()) @ (b.double().T if tb else b.double())
        try:
            c = ops.gemm(ctx, a.cuda(), b.cuda(), trans_a=ta, trans_b=tb, math=ZB_MATH_TF32)
            torch.cuda.synchronize()
            ctx.check()
            got = c.cpu().double()
            err = (got - ref).abs()
            scale = ref.abs().max().item()
            rel = (err.norm() / ref.norm()).item()
            bad = []
            for mb in range((m + 127) // 128):
                for nb in range((n + 31) // 32):
                    e = err[mb * 128:(mb + 1) * 128, nb * 32:(nb + 1) * 32].max().item()
                    if e > 2e-2 * scale:
                        bad.append((mb, nb))
            print(json.dumps({"m": m, "n": n, "k": k, "ta": ta, "tb": tb, "rel": round(rel, 6), "bad_blocks": len(bad),
                              "first_bad": bad[:12]}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"m": m, "n": n, "k": k, "ta": ta, "tb": tb, "error": repr(e)[:200]}), flush=True)
            ctx = ops.Context(math=ZB_MATH_TF32)


if __name__ == "__main__":
    main()
