#!/usr/bin/env python3
"""Experiment: can a K-major SWIZZLE_128B UMMA descriptor start at a 128-byte row that is not 1024-byte aligned
(needed to reuse one smem halo tile for all filter taps)?  D[i] should equal A[i + shift] . B^T for i < 128 - shift."""
import os
import subprocess
import sys

if len(sys.argv) > 1:
    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from zenu_b200 import ops
    shift = int(os.environ["ZENU_B200_DBG_ASHIFT"].split(",")[0])
    ctx = ops.Context()
    rng = np.random.default_rng(0)
    a = rng.standard_normal((128, 64)).astype(np.float32)
    b = rng.standard_normal((64, 64)).astype(np.float32)
    c = ops.gemm(ctx, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), trans_b=True).cpu().numpy()
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    n = 128 - shift
    err = np.abs(c[:n] - ref[shift:shift + n]).max() / np.abs(ref).max()
    err_unshifted = np.abs(c[:n] - ref[:n]).max() / np.abs(ref).max()
    print(f"shift/mode {os.environ['ZENU_B200_DBG_ASHIFT']}: err vs shifted rows {err:.3e}   (vs unshifted {err_unshifted:.3e})")
    ctx.close()
else:
    for mode in (1, 2):
        for shift in (0, 1, 2, 3, 7, 8, 9, 58):
            env = dict(os.environ, ZENU_B200_DBG_ASHIFT=f"{shift},{mode}")
            subprocess.run([sys.executable, __file__, "run"], env=env)
