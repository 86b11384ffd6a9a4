#!/usr/bin/env python3
"""Per-node timing of one train step (CUDA events around every tape node, zb_model_profile_*): where the step goes.
GPU box only.  Usage: python tools/profile_step.py [--arch resnet50] [--batch 256] [--steps 2] [--out gpurun_out/step_profile.tsv]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenu_b200 import nn, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="resnet50")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--hw", type=int, default=224)
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--out", default="gpurun_out/step_profile.tsv")
    a = ap.parse_args()
    ctx = ops.Context()
    model = nn.Model(ctx, a.arch, a.classes, fused=True, seed=42)
    model.set_optimizer("sgd", lr=0.01)
    g = torch.Generator().manual_seed(1234)
    X = torch.randn((a.batch, 3, a.hw, a.hw), generator=g).cuda()
    T = torch.zeros((a.batch, a.classes))
    T[torch.arange(a.batch), torch.randint(0, a.classes, (a.batch,), generator=g)] = 1.0
    T = T.cuda()
    for _ in range(3):
        model.train_step(X, T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        model.train_step(X, T)
    e1.record()
    torch.cuda.synchronize()
    plain_ms = e0.elapsed_time(e1) / a.steps
    model.profile(True)
    e0.record()
    for _ in range(a.steps):
        model.train_step(X, T)
    e1.record()
    torch.cuda.synchronize()
    prof_ms = e0.elapsed_time(e1) / a.steps
    rows = model.profile_table()
    model.profile(False)
    tot = sum(r[2] for r in rows) / a.steps
    rows.sort(key=lambda r: -r[2])
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        hdr = f"# {a.arch} batch {a.batch}: step {plain_ms:.2f} ms (unprofiled), {prof_ms:.2f} ms (profiled), sum of nodes {tot:.2f} ms/step"
        print(hdr)
        f.write(hdr + "\n# key\tcount/step\tms/step\tms/call\tTFLOP/s\tGB/s\tshare\n")
        for k, n, ms, fl, by in rows:
            per = ms / n
            line = f"{k}\t{n / a.steps:g}\t{ms / a.steps:.3f}\t{per:.4f}\t{fl / ms / 1e9 if fl else 0:.1f}\t{by / ms / 1e6 if by else 0:.0f}\t{ms / a.steps / tot * 100:.1f}%"
            f.write(line + "\n")
        # per class
        cls = {}
        for k, n, ms, fl, by in rows:
            c = k.split(" ")[0]
            cls.setdefault(c, [0.0, 0.0, 0.0])
            cls[c][0] += ms / a.steps
            cls[c][1] += fl / a.steps
            cls[c][2] += by / a.steps
        f.write("# ---- by class\n")
        for c, (ms, fl, by) in sorted(cls.items(), key=lambda kv: -kv[1][0]):
            line = f"# {c}\t{ms:.3f} ms/step\t{fl / ms / 1e9 if fl else 0:.1f} TFLOP/s\t{by / ms / 1e6 if by else 0:.0f} GB/s\t{ms / tot * 100:.1f}%"
            print(line)
            f.write(line + "\n")
    for r in rows[:25]:
        print(f"{r[0]:60s} n={r[1] / a.steps:g} {r[2] / a.steps:9.3f} ms/step")
    model.close()
    ctx.close()


if __name__ == "__main__":
    main()
