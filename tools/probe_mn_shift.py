#!/usr/bin/env python3
"""Experiment: MN-major (SWIZZLE_128B_BASE32B) UMMA B descriptors whose start is a 128-byte row that is not 512/1024-byte
aligned, and whose N boxes overlap (LBO = 128 B: box j = box 0 shifted by j rows).  Needed to serve all filter taps of a
wgrad from ONE smem raster of X (K = pixels, N = channels).  C = A[m,k] . B[k,n] through ops.gemm (B stored [k][n])."""
import os
import subprocess
import sys

if len(sys.argv) > 1:
    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from zenu_b200 import ops
    shift, lbo = [int(v) for v in os.environ["ZENU_B200_DBG_BSHIFT"].split(",")]
    ctx = ops.Context()
    rng = np.random.default_rng(0)
    if lbo == 0:
        # row shift: the box is loaded `shift` rows early (rows < 0 zero-filled), the descriptor starts `shift` rows later;
        # K = 32 - shift so the rows read beyond the box only meet zero-filled A columns
        k = 32 - shift
        a = np.zeros((128, 32), np.float32); a[:, :k] = rng.standard_normal((128, k))
        b = np.zeros((32, 32), np.float32); b[:k] = rng.standard_normal((k, 32))
        A, B = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        c = ops.gemm(ctx, A, B, trans_b=False).cpu().numpy()
        ref = a.astype(np.float64) @ b.astype(np.float64)
        print(f"B row shift {shift}: max err {np.abs(c - ref).max() / np.abs(ref).max():.3e}")
    else:
        # overlapped N boxes: columns 32..63 of C = A . (box 0 shifted down by one row; its row 31 is row 0 of box 1)
        a = rng.standard_normal((128, 32)).astype(np.float32)
        b = rng.standard_normal((32, 64)).astype(np.float32)
        c = ops.gemm(ctx, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), trans_b=False).cpu().numpy()
        b2 = np.concatenate([b[1:32, :32], b[0:1, 32:64]], 0)
        ref0 = a.astype(np.float64) @ b[:, :32].astype(np.float64)
        ref1 = a.astype(np.float64) @ b2.astype(np.float64)
        print(f"LBO {lbo}: box0 err {np.abs(c[:, :32] - ref0).max() / np.abs(ref0).max():.3e}  box1 (shifted view) err "
              f"{np.abs(c[:, 32:] - ref1).max() / np.abs(ref1).max():.3e}")
    ctx.close()
else:
    for env_v in ("0,0", "1,0", "2,0", "3,0", "4,0", "5,0", "9,0", "0,128"):
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, ZENU_B200_DBG_BSHIFT=env_v))
