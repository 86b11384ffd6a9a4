#!/usr/bin/env python3
"""One warm-up + N train steps of the bench workload with nothing else around it: the command ncu wraps for the launch
list and the --set full capture (B200_PROFILING.md).  Usage: python tools/ncu_step.py [--steps 1] [--batch 256]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenu_b200 import nn, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--arch", default="resnet50")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--hw", type=int, default=224)
ap.add_argument("--classes", type=int, default=1000)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
a = ap.parse_args()
ctx = ops.Context()
model = nn.Model(ctx, a.arch, a.classes, fused=True, seed=42)
model.set_optimizer("sgd", lr=0.01)
g = torch.Generator().manual_seed(1234)
X = torch.randn((a.batch, 3, a.hw, a.hw), generator=g).cuda()
T = torch.zeros((a.batch, a.classes))
T[torch.arange(a.batch), torch.randint(0, a.classes, (a.batch,), generator=g)] = 1.0
T = T.cuda()
for _ in range(a.warmup):
    model.train_step(X, T)
torch.cuda.synchronize()
n0 = ctx.launch_count()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.steps):
    model.train_step(X, T)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("launches per step:", (ctx.launch_count() - n0) / a.steps)
model.close()
ctx.close()
