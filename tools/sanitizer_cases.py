#!/usr/bin/env python3
"""One small invocation of every kernel family of the hot path, for compute-sanitizer (memcheck / racecheck / synccheck run this,
tools/sanitize.sh).  Each case is also checked against the oracle so a sanitizer run that perturbs timing still proves correctness.
Shapes are chosen to hit: the implicit-GEMM umma_kernel (1x1, strided, split-K, beta epilogue), CTA pairs, the halo conv kernel
(resident / streamed filter, pairs), wgrad_halo, the three stem kernels, BatchNorm reduce / finalize / apply (fused ReLU / residual /
bit mask), pooling, the loss head, the optimizers and the layout kernels."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import zenu_oracle as zo  # noqa: E402
from zenu_b200 import ZB_MATH_TF32, ZB_MATH_TF32X3, ZB_NHWC, nn, ops  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def nhwc(a):
    return np.ascontiguousarray(np.transpose(a, (0, 2, 3, 1)))


def nchw(t):
    return np.ascontiguousarray(np.transpose(t.detach().cpu().numpy(), (0, 3, 1, 2)))


def rel(a, b):
    return float(np.linalg.norm((np.asarray(a, np.float64) - b).ravel()) / (np.linalg.norm(np.asarray(b, np.float64).ravel()) + 1e-30))


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    ctx = ops.Context()
    rng = np.random.default_rng(0)
    zo.use_openblas()
    convs = [  # n, c, h, k, r, pad, stride
        ("umma 1x1", 2, 64, 14, 128, 1, 0, 1), ("umma 1x1 pair", 4, 512, 16, 256, 1, 0, 1), ("umma strided 3x3", 2, 64, 15, 64, 3, 1, 2),
        ("umma 1x1 s2", 2, 64, 16, 96, 1, 0, 2), ("halo resident", 2, 64, 18, 64, 3, 1, 1), ("halo streamed pair", 2, 128, 14, 256, 3, 1, 1),
        ("stem 7x7", 2, 3, 40, 64, 7, 3, 2), ("stem 3x3 c4", 1, 4, 30, 64, 3, 1, 1), ("small 7x7 stage", 4, 256, 7, 256, 3, 1, 1),
    ]
    for name, n, c, h, k, r, pad, stride in convs:
        if only and only not in name:
            continue
        x = rng.standard_normal((n, c, h, h)).astype(np.float32)
        w = (rng.standard_normal((k, c, r, r)) * np.sqrt(2.0 / (c * r * r))).astype(np.float32)
        y_ref = zo.conv2d_fwd(x.astype(np.float64), w.astype(np.float64), pad, stride, 1)
        dy = rng.standard_normal(y_ref.shape).astype(np.float32)
        X, W, DY = dev(nhwc(x)), dev(nhwc(w)), dev(nhwc(dy))
        shift = dev(np.zeros(k, np.float32))
        for math, tol in ((ZB_MATH_TF32, 2e-3), (ZB_MATH_TF32X3, 2e-5)):
            kw = dict(pad=pad, stride=stride, dil=1, layout=ZB_NHWC, math=math)
            y, _, _ = ops.conv_fwd_bnstats(ctx, X, W, shift, **kw)
            dx = ops.conv_bkwd_data(ctx, DY, W, X.shape, **kw)
            dw = ops.conv_bkwd_weight(ctx, DY, X, W.shape, **kw)
            acc = torch.ones_like(dx)
            if c % 32 == 0:
                ops.conv_bkwd_data_accumulate(ctx, DY, W, acc, **kw)
                words = torch.randint(-2 ** 31, 2 ** 31 - 1, ((acc.numel() + 31) // 32,), dtype=torch.int64, device="cuda").to(torch.int32)
                want = ops.mask_apply(ctx, acc, words)
                ops.conv_bkwd_data_accumulate(ctx, DY, W, want, **kw)
                ops.conv_bkwd_data_accumulate_masked(ctx, DY, W, acc, words, **kw)   # lazy masked fan-in (epilogue mask prefetch)
                assert torch.equal(acc, want), (name, "masked accumulate")
            ctx.check()
            assert rel(nchw(y), y_ref) < tol, (name, "fprop")
            assert rel(nchw(dx), zo.conv2d_bkwd_data(dy.astype(np.float64), w.astype(np.float64), x.shape, pad, stride, 1)) < tol, (name, "dgrad")
            assert rel(nchw(dw), zo.conv2d_bkwd_filter(dy.astype(np.float64), x.astype(np.float64), w.shape, pad, stride, 1)) < tol, (name, "wgrad")
        print("ok conv", name, flush=True)
    if not only or only in "gemm":
        for (m, n, k) in ((300, 200, 100), (512, 256, 1024), (8, 8, 4096)):
            a, b = rng.standard_normal((m, k)).astype(np.float32), rng.standard_normal((k, n)).astype(np.float32)
            c0 = rng.standard_normal((m, n)).astype(np.float32)
            got = ops.gemm(ctx, dev(a), dev(b), False, False, 0.5, 0.25, dev(c0))
            assert rel(got.cpu().numpy(), zo.gemm(a.astype(np.float64), b.astype(np.float64), False, False, 0.5, 0.25, c0.astype(np.float64))) < 2e-3
        print("ok gemm", flush=True)
    if not only or only in "bn":
        for shape in ((4, 64, 9, 7), (2, 256, 6, 6), (3, 5, 6, 6)):
            n, c, h, w = shape
            x = rng.standard_normal(shape).astype(np.float32)
            res = rng.standard_normal(shape).astype(np.float32)
            dy = rng.standard_normal(shape).astype(np.float32)
            sc, bi = rng.uniform(0.5, 1.5, c).astype(np.float32), rng.standard_normal(c).astype(np.float32)
            X, R, DY = dev(nhwc(x)), dev(nhwc(res)), dev(nhwc(dy))
            rm, rv = dev(np.zeros(c, np.float32)), dev(np.ones(c, np.float32))
            y, sm, si = ops.batch_norm_2d_forward_train(ctx, 0.9, X, dev(sc), dev(bi), rm, rv, layout=ZB_NHWC, residual=R, relu=True)
            dx, ds, db, dres = ops.batch_norm_2d_backward(ctx, X, DY, dev(sc), sm, si, layout=ZB_NHWC, y=y, want_residual_grad=True)
            dx2, _, _ = ops.batch_norm_2d_relu_backward(ctx, X, DY, dev(sc), dev(bi), sm, si, layout=ZB_NHWC)
            if c % 32 == 0:
                y3, sm3, si3, mask = ops.batch_norm_2d_forward_train_masked(ctx, 0.9, X, dev(sc), dev(bi), dev(np.zeros(c, np.float32)),
                                                                            dev(np.ones(c, np.float32)), residual=R)
                ops.batch_norm_2d_backward_masked(ctx, X, DY, dev(sc), sm3, si3, mask)
                ops.batch_norm_2d_backward_masked(ctx, X, DY, dev(sc), sm3, si3, mask, want_residual_grad=False)
            if c % 4 == 0 and (256 % (c // 4)) == 0:   # fused stem: BN + ReLU + 3x3 / 2 max-pool, both ways
                yp, pidx, sm4, si4 = ops.batch_norm_relu_max_pool_forward_train(ctx, 0.9, X, dev(sc), dev(bi), dev(np.zeros(c, np.float32)),
                                                                                dev(np.ones(c, np.float32)))
                ops.batch_norm_relu_max_pool_backward(ctx, X, torch.randn_like(yp), pidx, dev(sc), dev(bi), sm4, si4)
            ctx.check()
            bn, _, _, sm_r, si_r = zo.bn2d_fwd_train(x, sc, bi, np.zeros(c, np.float32), np.ones(c, np.float32), 0.9)
            out_ref = zo.relu(zo.ewise("add", bn, res))
            assert rel(nchw(y), out_ref) < 1e-5
            g = zo.ewise("mul", dy, (out_ref > 0).astype(np.float32))
            assert rel(nchw(dx), zo.bn2d_bwd(x, g, sc, sm_r, si_r)[0]) < 1e-3
        print("ok bn", flush=True)
    if not only or only in "pool loss optim layout":
        x = np.maximum(rng.standard_normal((2, 8, 11, 9)), 0).astype(np.float32)
        y, idx = ops.max_pool_2d_indexed(ctx, dev(nhwc(x)), 3, 2, 1)
        assert np.array_equal(nchw(y), zo.maxpool2d_fwd(x, 3, 2, 1))
        dyp = rng.standard_normal(tuple(y.shape)).astype(np.float32)
        ops.max_pool_2d_indexed_backward(ctx, dev(dyp), idx, nhwc(x).shape, 3, 2, 1)
        ops.global_avg_pool(ctx, dev(nhwc(x)), layout=ZB_NHWC)
        z = (rng.standard_normal((16, 1000)) * 3).astype(np.float32)
        t = np.zeros((16, 1000), np.float32)
        t[np.arange(16), rng.integers(0, 1000, 16)] = 1.0
        loss, dz = ops.softmax_cross_entropy(ctx, dev(z), dev(t))
        assert abs(float(loss.cpu()[0]) - zo.softmax_xent(z.astype(np.float64), t.astype(np.float64))[0]) < 1e-4
        p, g = dev(rng.standard_normal(10007).astype(np.float32)), dev(rng.standard_normal(10007).astype(np.float32))
        ops.sgd_step(ctx, p, g, 0.01)
        ops.adam_step(ctx, p, g, torch.zeros_like(p), torch.zeros_like(p), 0.01, 0.9, 0.999, 1e-8, 1, 0.01, True)
        ops.to_nchw(ctx, ops.to_nhwc(ctx, dev(x)))
        ops.sum_to(ctx, dev(x), (8, 1, 1))
        ops.variance_axis(ctx, dev(x), 1)
        ctx.check()
        print("ok pool / loss / optimizers / layout / reductions", flush=True)
    if not only or only in "model":
        model = nn.Model(ctx, "resnet18", 10, seed=1)
        model.set_optimizer("sgd", lr=1e-3)
        X = dev(rng.standard_normal((4, 3, 64, 64)).astype(np.float32))
        T = dev(np.eye(10, dtype=np.float32)[rng.integers(0, 10, 4)])
        losses = [model.train_step(X, T, read_loss=True) for _ in range(2)]
        assert all(np.isfinite(losses))
        ctx.check()
        model.close()
        print("ok resnet18 train steps", losses, flush=True)
    ctx.close()
    print("SANITIZER_CASES_DONE", flush=True)


if __name__ == "__main__":
    main()
