#!/usr/bin/env python3
"""CPU baseline table of BASELINE.md section 4: the reference's CPU algorithm (oracle port, OpenBLAS on all host cores for the GEMMs,
everything else single-threaded exactly as the Rust is) timed on the GPU box's host cores.  Baseline only -- no target attached.
  cfg1 small CNN train step in full (N = 64); ResNet-18 / ResNet-50 train steps at N = 8 (per image, extrapolates linearly);
  conv fwd / dgrad / wgrad for every unique ResNet-50 shape at N = 8; BatchNorm(+ReLU) fwd / bwd effective GB/s.
Usage: python tools/cpu_baseline.py [--out gpurun_out/cpu_baseline.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import zenu_oracle as zo  # noqa: E402
from oracle import zenu_oracle_model as zm  # noqa: E402

SHAPES = [(3, 224, 64, 7, 2, 3), (64, 56, 64, 1, 1, 0), (64, 56, 64, 3, 1, 1), (64, 56, 256, 1, 1, 0), (256, 56, 64, 1, 1, 0),
          (256, 56, 128, 1, 1, 0), (128, 56, 128, 3, 2, 1), (128, 28, 512, 1, 1, 0), (256, 56, 512, 1, 2, 0), (512, 28, 128, 1, 1, 0),
          (128, 28, 128, 3, 1, 1), (512, 28, 256, 1, 1, 0), (256, 28, 256, 3, 2, 1), (256, 14, 1024, 1, 1, 0), (512, 28, 1024, 1, 2, 0),
          (1024, 14, 256, 1, 1, 0), (256, 14, 256, 3, 1, 1), (1024, 14, 512, 1, 1, 0), (512, 14, 512, 3, 2, 1), (512, 7, 2048, 1, 1, 0),
          (1024, 14, 2048, 1, 2, 0), (2048, 7, 512, 1, 1, 0), (512, 7, 512, 3, 1, 1)]


def timed(f, reps=2):
    f()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    return (time.perf_counter() - t0) / reps


def train(arch, n, hw, classes, steps):
    m = zm.OracleModel(arch, classes, zm.init_params(arch, classes, seed=42))
    rng = np.random.default_rng(1)
    x = rng.standard_normal((n, 3, hw, hw)).astype(np.float32)
    t = np.zeros((n, classes), np.float32)
    t[np.arange(n), rng.integers(0, classes, n)] = 1.0
    sec = timed(lambda: m.train_step(x, t, kind="sgd", lr=0.01), steps)
    return {"batch": n, "sec_per_step": sec, "images_per_s": n / sec}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/cpu_baseline.json")
    a = ap.parse_args()
    cores = os.cpu_count() or 1
    blas = zo.use_openblas(threads=cores)
    out = {"cores": cores, "blas": "OpenBLAS (numpy bundled), all cores" if blas else "plain C loops", "kind": "port"}
    try:
        out["cpu_model"] = [ln.split(":")[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")][0]
    except Exception:  # noqa: BLE001
        pass
    out["small_cnn_n64"] = train("small_cnn", 64, 32, 10, 5)
    out["resnet18_n8"] = train("resnet18", 8, 224, 1000, 2)
    out["resnet50_n8"] = train("resnet50", 8, 224, 1000, 2)
    rng = np.random.default_rng(2)
    n = 8
    conv = []
    for ci, h, co, k, s, p in SHAPES:
        ho = (h + 2 * p - k) // s + 1
        x = rng.standard_normal((n, ci, h, h)).astype(np.float32)
        w = rng.standard_normal((co, ci, k, k)).astype(np.float32)
        dy = rng.standard_normal((n, co, ho, ho)).astype(np.float32)
        gflop = 2.0 * n * co * ho * ho * ci * k * k / 1e9
        tf = timed(lambda: zo.conv2d_fwd(x, w, p, s, 1))
        td = timed(lambda: zo.conv2d_bkwd_data(dy, w, x.shape, p, s, 1))
        tw = timed(lambda: zo.conv2d_bkwd_filter(dy, x, w.shape, p, s, 1))
        conv.append({"shape": f"{ci}x{h}->{co} k{k} s{s}", "n": n, "gflop": gflop, "fwd_ms": tf * 1e3, "dgrad_ms": td * 1e3, "wgrad_ms": tw * 1e3,
                     "fwd_gflops": gflop / tf, "dgrad_gflops": gflop / td, "wgrad_gflops": gflop / tw})
    out["conv_n8"] = conv
    bn = []
    for c, h in ((64, 112), (256, 56), (512, 28), (1024, 14), (2048, 7)):
        x = rng.standard_normal((n, c, h, h)).astype(np.float32)
        dy = rng.standard_normal((n, c, h, h)).astype(np.float32)
        sc, bi = np.ones(c, np.float32), np.zeros(c, np.float32)
        y, _, _, sm, si = zo.bn2d_fwd_train(x, sc, bi, np.zeros(c, np.float32), np.ones(c, np.float32), 0.9)
        tf = timed(lambda: zo.relu(zo.bn2d_fwd_train(x, sc, bi, np.zeros(c, np.float32), np.ones(c, np.float32), 0.9)[0]))
        tb = timed(lambda: zo.bn2d_bwd(x, zo.ewise("mul", dy, zo.relu_backward_mask(y)), sc, sm, si))
        mb = x.nbytes / 1e6
        bn.append({"shape": f"{n}x{c}x{h}x{h}", "mb": mb, "fwd_relu_ms": tf * 1e3, "bwd_relu_ms": tb * 1e3,
                   "fwd_relu_gbs_on_12B_per_elem": 3 * mb / tf / 1e3, "bwd_relu_gbs_on_20B_per_elem": 5 * mb / tb / 1e3})
    out["bn_relu_n8"] = bn
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if not isinstance(v, list)}))


if __name__ == "__main__":
    main()
