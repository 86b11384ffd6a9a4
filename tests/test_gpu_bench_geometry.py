"""Parity at the geometry the benchmark runs (VERDICT r1 "what's weak" 1): the 23 unique conv shapes of ResNet-50 (SURVEY Table
8d-1) at their REAL spatial sizes (224 stem; 56 / 28 / 14 / 7) with a batch that gives ~50 k GEMM rows (several waves of 128-row
tiles on 148 SMs, ring wraps, CTA pairs, split-K), fprop / dgrad / wgrad through the C ABI against the CPU oracle:

  * TF32 vs the oracle fed with tf32-rounded operands at 5e-5 (only the fp32 summation order differs, so a kernel bug cannot hide
    inside the 1e-3 TF32 allowance), 3xTF32 vs the oracle's exact arithmetic (f64) at 1e-5;
  * `zb_conv2d_plan_describe` (a dry run of the same host planners) must report, for the tested batch, the SAME kernel variant
    (kernel, tile width, pipeline depth, CTA pairs, halo / resident filter, split-K, fused statistics ...) it reports for the
    benchmarked batch of 256, and the trace of the real call must equal the dry run: the test exercises what the bench exercises;
  * BatchNorm forward / backward at 256 x 256 x 56 x 56 (822 MB per tensor: the 32-bit offset paths) against the f64 oracle.

Reference contract: zenu-matrix/src/nn/conv/mod.rs:131-179 (tolerances), zenu-test/src/lib.rs:3-24 (max-abs-diff comparison).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import zenu_oracle as zo  # noqa: E402

pytestmark = pytest.mark.gpu

# SURVEY Table 8d-1: (#, C_in, H_in, C_out, k, stride, pad)
RESNET50_SHAPES = [
    (1, 3, 224, 64, 7, 2, 3), (2, 64, 56, 64, 1, 1, 0), (3, 64, 56, 64, 3, 1, 1), (4, 64, 56, 256, 1, 1, 0),
    (5, 256, 56, 64, 1, 1, 0), (6, 256, 56, 128, 1, 1, 0), (7, 128, 56, 128, 3, 2, 1), (8, 128, 28, 512, 1, 1, 0),
    (9, 256, 56, 512, 1, 2, 0), (10, 512, 28, 128, 1, 1, 0), (11, 128, 28, 128, 3, 1, 1), (12, 512, 28, 256, 1, 1, 0),
    (13, 256, 28, 256, 3, 2, 1), (14, 256, 14, 1024, 1, 1, 0), (15, 512, 28, 1024, 1, 2, 0), (16, 1024, 14, 256, 1, 1, 0),
    (17, 256, 14, 256, 3, 1, 1), (18, 1024, 14, 512, 1, 1, 0), (19, 512, 14, 512, 3, 2, 1), (20, 512, 7, 2048, 1, 1, 0),
    (21, 1024, 14, 2048, 1, 2, 0), (22, 2048, 7, 512, 1, 1, 0), (23, 512, 7, 512, 3, 1, 1),
]
BENCH_BATCH = 256


def candidate_batches(h_out):
    """Batches that give >= ~50 k output pixels (or the benchmarked batch itself where that is smaller)."""
    want = max(16, -(-50176 // (h_out * h_out)))
    cands = [want, want + want // 2, 2 * want, want + 8, BENCH_BATCH]
    return [n for i, n in enumerate(cands) if n <= BENCH_BATCH and n not in cands[:i]]


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import zenu_b200
    from zenu_b200 import ops
    ctx = ops.Context()
    zo.use_openblas()
    yield zenu_b200, ops, ctx
    zo.use_plain_gemm()
    ctx.check()
    ctx.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def nhwc(a):
    return np.ascontiguousarray(np.transpose(a, (0, 2, 3, 1)))


def nchw(a):
    return np.ascontiguousarray(np.transpose(a, (0, 3, 1, 2)))


def rel_err(got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    rel_l2 = np.linalg.norm((got - ref).ravel()) / (np.linalg.norm(ref.ravel()) + 1e-300)
    max_rel = np.max(np.abs(got - ref)) / (np.max(np.abs(ref)) + 1e-300)
    return max(rel_l2, max_rel)


def plans(ops, pkg, ctx, n, ci, h, co, k, stride, pad, math):
    xs, ws = (n, h, h, ci), (co, k, k, ci)
    kw = dict(pad=pad, stride=stride, dil=1, layout=pkg.ZB_NHWC, math=math)
    return {
        "fprop+bnstats": ops.conv_plan_describe(ctx, ops.PLAN_FPROP, xs, ws, flags=ops.PLAN_BNSTATS, **kw),
        "dgrad": ops.conv_plan_describe(ctx, ops.PLAN_DGRAD, xs, ws, **kw),
        "wgrad": ops.conv_plan_describe(ctx, ops.PLAN_WGRAD, xs, ws, **kw),
    }


@pytest.mark.parametrize("shape", RESNET50_SHAPES, ids=lambda s: f"L{s[0]}_{s[1]}x{s[2]}to{s[3]}_k{s[4]}s{s[5]}")
def test_resnet50_conv_shape_at_bench_geometry(env, shape):
    pkg, ops, ctx = env
    _, ci, h, co, k, stride, pad = shape
    h_out = (h + 2 * pad - k) // stride + 1
    tf32 = pkg.ZB_MATH_TF32
    want = {op: ops.plan_variant(t) for op, t in plans(ops, pkg, ctx, BENCH_BATCH, ci, h, co, k, stride, pad, tf32).items()}
    n = None
    for cand in candidate_batches(h_out):
        got = {op: ops.plan_variant(t) for op, t in plans(ops, pkg, ctx, cand, ci, h, co, k, stride, pad, tf32).items()}
        if got == want:
            n = cand
            break
    assert n is not None, f"no test batch reproduces the kernel variants of batch {BENCH_BATCH}: {want}"
    dry = plans(ops, pkg, ctx, n, ci, h, co, k, stride, pad, tf32)

    rng = np.random.default_rng(1000 + shape[0])
    x = rng.standard_normal((n, ci, h, h), dtype=np.float32)
    w = (rng.standard_normal((co, ci, k, k)) * np.sqrt(2.0 / (ci * k * k))).astype(np.float32)
    dy = rng.standard_normal((n, co, h_out, h_out), dtype=np.float32)
    X, W, DY = dev(nhwc(x)), dev(nhwc(w)), dev(nhwc(dy))
    shift = dev((0.1 * rng.standard_normal(co)).astype(np.float32))
    kw = dict(pad=pad, stride=stride, dil=1, layout=pkg.ZB_NHWC)

    # ---- TF32: the real calls, traced; must be the launches the dry run described
    ops.plan_trace(ctx, True)
    y, partial, rows = ops.conv_fwd_bnstats(ctx, X, W, shift, math=tf32, **kw)
    t_fprop = ops.plan_trace_read(ctx)
    dx = ops.conv_bkwd_data(ctx, DY, W, X.shape, math=tf32, **kw)
    t_dgrad = ops.plan_trace_read(ctx)
    dw = ops.conv_bkwd_weight(ctx, DY, X, W.shape, math=tf32, **kw)
    t_wgrad = ops.plan_trace_read(ctx)
    ops.plan_trace(ctx, False)
    ctx.check()
    assert (t_fprop, t_dgrad, t_wgrad) == (dry["fprop+bnstats"], dry["dgrad"], dry["wgrad"])
    assert rows > 0, "the bench's fused BatchNorm statistics were not taken for this shape"

    xr, wr, dyr = zo.tf32_round(x, "rne"), zo.tf32_round(w, "rne"), zo.tf32_round(dy, "rne")
    y_ref = zo.conv2d_fwd(xr, wr, pad, stride, 1)
    y_h = nchw(host(y))
    assert rel_err(y_h, y_ref) < 5e-5, "fprop"
    # fused statistics: per-channel sum(y - shift), sum((y - shift)^2) of the values the kernel stored
    part = host(partial)[:rows].astype(np.float64).sum(axis=0)
    d = y_h.astype(np.float64) - host(shift).astype(np.float64)[None, :, None, None]
    np.testing.assert_allclose(part[0], d.sum(axis=(0, 2, 3)), rtol=1e-4, atol=1e-3 * np.sqrt(d[:, 0].size))
    np.testing.assert_allclose(part[1], (d * d).sum(axis=(0, 2, 3)), rtol=1e-4)
    del d, y_h
    assert rel_err(nchw(host(dx)), zo.conv2d_bkwd_data(dyr, wr, x.shape, pad, stride, 1)) < 5e-5, "dgrad"
    assert rel_err(nchw(host(dw)), zo.conv2d_bkwd_filter(dyr, xr, w.shape, pad, stride, 1)) < 5e-5, "wgrad"
    # gradient fan-in form the residual blocks use (dx += dgrad), where the tensor-core path serves it
    if ci % 32 == 0:
        acc = torch.ones_like(dx)
        ops.conv_bkwd_data_accumulate(ctx, DY, W, acc, math=tf32, **kw)
        assert rel_err(host(acc) - 1.0, host(dx)) < 1e-4, "dgrad accumulate"
        del acc
        # ... and with a lazily masked first arrival (the residual gradient of a fused BN + add + ReLU): bit-identical to
        # accumulating onto the materialised product old (.) mask, whichever path serves the shape
        old = torch.randn_like(dx)
        words = torch.randint(-2 ** 31, 2 ** 31 - 1, ((old.numel() + 31) // 32,), dtype=torch.int64, device="cuda").to(torch.int32)
        want_acc = ops.mask_apply(ctx, old, words)
        bits = np.unpackbits(host(words).view(np.uint8), bitorder="little")[: old.numel()].astype(bool)
        np.testing.assert_array_equal(host(want_acc).ravel(), np.where(bits, host(old).ravel(), np.float32(0)))
        ops.conv_bkwd_data_accumulate(ctx, DY, W, want_acc, math=tf32, **kw)
        ops.conv_bkwd_data_accumulate_masked(ctx, DY, W, old, words, math=tf32, **kw)
        np.testing.assert_array_equal(host(old), host(want_acc))
        del old, words, want_acc
    del y, dx, dw, xr, wr, dyr, y_ref

    # ---- 3xTF32 against exact arithmetic
    x3 = pkg.ZB_MATH_TF32X3
    x64, w64, dy64 = x.astype(np.float64), w.astype(np.float64), dy.astype(np.float64)
    y = ops.conv_fwd(ctx, X, W, math=x3, **kw)
    assert rel_err(nchw(host(y)), zo.conv2d_fwd(x64, w64, pad, stride, 1)) < 1e-5, "fprop 3xTF32"
    del y
    dx = ops.conv_bkwd_data(ctx, DY, W, X.shape, math=x3, **kw)
    assert rel_err(nchw(host(dx)), zo.conv2d_bkwd_data(dy64, w64, x.shape, pad, stride, 1)) < 1e-5, "dgrad 3xTF32"
    del dx
    dw = ops.conv_bkwd_weight(ctx, DY, X, W.shape, math=x3, **kw)
    assert rel_err(nchw(host(dw)), zo.conv2d_bkwd_filter(dy64, x64, w.shape, pad, stride, 1)) < 1e-5, "wgrad 3xTF32"
    ctx.check()


def test_plan_describe_reports_every_kernel_family(env):
    """The description names the kernel families DESIGN.md says serve the ResNet-50 layers at batch 256 (a planner regression that
    silently sent a layer to a slower family would change these strings)."""
    pkg, ops, ctx = env
    tf32 = pkg.ZB_MATH_TF32

    def p(op, ci, h, co, k, stride, pad, flags=0):
        return ops.conv_plan_describe(ctx, op, (BENCH_BATCH, h, h, ci), (co, k, k, ci), pad=pad, stride=stride, dil=1,
                                      layout=pkg.ZB_NHWC, math=tf32, flags=flags)

    assert p(ops.PLAN_FPROP, 3, 224, 64, 7, 2, 3).count("stem_fprop<bn=64>") == 1
    assert "stem_dgrad " in p(ops.PLAN_DGRAD, 3, 224, 64, 7, 2, 3)
    assert "stem_wgrad " in p(ops.PLAN_WGRAD, 3, 224, 64, 7, 2, 3)
    assert "halo_conv<bn=64,cl=2,pair=1> resident=1" in p(ops.PLAN_FPROP, 64, 56, 64, 3, 1, 1)
    assert "halo_conv<bn=256,cl=2,pair=1>" in p(ops.PLAN_DGRAD, 256, 14, 256, 3, 1, 1)
    assert "wgrad_halo " in p(ops.PLAN_WGRAD, 128, 28, 128, 3, 1, 1)
    assert ",cl=2>" in p(ops.PLAN_FPROP, 1024, 14, 512, 1, 1, 0)                       # CTA pairs on the wide, long-K pointwise layers
    assert "splitk=1" in p(ops.PLAN_WGRAD, 512, 7, 2048, 1, 1, 0)                      # split-K wgrad + deterministic reduce
    d_s2 = p(ops.PLAN_DGRAD, 128, 56, 128, 3, 2, 1)                                    # one launch per (h, w) parity class:
    assert d_s2.count("umma<") == 1 and d_s2.count("halo_conv<") == 3                  # the single-tap class as a GEMM, the 2- / 2- / 4-tap classes on the halo kernel
    assert "planes=4" in p(ops.PLAN_FPROP, 128, 56, 128, 3, 2, 1)                      # stride-2 forward: four input parity planes per raster slot
    assert "simt_conv_fprop<f64>" in ops.conv_plan_describe(ctx, ops.PLAN_FPROP, (2, 8, 8, 32), (32, 3, 3, 32), pad=1,
                                                             layout=pkg.ZB_NHWC, dtype=torch.float64)
    with pytest.raises(pkg.ZenuB200Error):
        ops.conv_plan_describe(ctx, ops.PLAN_FPROP, (1, 4, 4, 32), (32, 9, 9, 32), pad=0, layout=pkg.ZB_NHWC)  # filter > input
    ctx.check()


def test_batchnorm_at_bench_tensor_size(env):
    """BatchNorm2d forward (train, + fused ReLU) and backward on the largest activation of the benchmarked step,
    256 x 256 x 56 x 56 f32 = 822 MB per tensor (element offsets beyond 2^27, byte offsets beyond 2^29), NHWC, against the f64 oracle
    (batch_norm.rs:283-391); tolerances as tests/test_gpu_parity.py::test_bn_vs_oracle."""
    pkg, ops, ctx = env
    n, c, h, w = 256, 256, 56, 56
    rng = np.random.default_rng(77)
    x = rng.standard_normal((n, c, h, w), dtype=np.float32)
    x *= 1.5
    x += 0.75
    dy = rng.standard_normal((n, c, h, w), dtype=np.float32)
    scale = rng.uniform(0.5, 1.5, c).astype(np.float32)
    bias = (0.2 * rng.standard_normal(c)).astype(np.float32)
    X, DY = dev(nhwc(x)), dev(nhwc(dy))
    rm, rv = dev(np.zeros(c, np.float32)), dev(np.ones(c, np.float32))
    y, sm, si = ops.batch_norm_2d_forward_train(ctx, 0.9, X, dev(scale), dev(bias), rm, rv, layout=pkg.ZB_NHWC)
    dx, ds, db = ops.batch_norm_2d_backward(ctx, X, DY, dev(scale), sm, si, layout=pkg.ZB_NHWC)
    ctx.check()
    y_h, dx_h = nchw(host(y)), nchw(host(dx))
    del y, dx, X, DY
    torch.cuda.empty_cache()
    x64, dy64 = x.astype(np.float64), dy.astype(np.float64)
    del x, dy
    f8 = lambda a: a.astype(np.float64)  # noqa: E731
    y_ref, rm_ref, rv_ref, sm_ref, si_ref = zo.bn2d_fwd_train(x64, f8(scale), f8(bias), np.zeros(c), np.ones(c), 0.9)
    tol = 2e-5
    assert rel_err(y_h, y_ref) < tol
    del y_ref, y_h
    assert rel_err(host(rm), rm_ref) < tol and rel_err(host(rv), rv_ref) < tol
    assert rel_err(host(sm), sm_ref) < tol and rel_err(host(si), si_ref) < tol
    dx_ref, ds_ref, db_ref = zo.bn2d_bwd(x64, dy64, f8(scale), sm_ref, si_ref)
    assert rel_err(dx_h, dx_ref) < 20 * tol
    assert rel_err(host(ds), ds_ref) < 20 * tol and rel_err(host(db), db_ref) < 20 * tol


def test_fused_stem_bn_relu_pool_at_bench_size(env):
    """zb_bn2d_relu_maxpool_fwd_train / _bwd on the stem activation of the benchmarked step (256 x 112 x 112 x 64 f32 = 822 MB, pooled to
    56 x 56): forward bit-identical to zb_bn2d_fwd_train(relu) + zb_maxpool2d_fwd_idx, backward equal to zb_maxpool2d_bwd_idx +
    zb_bn2d_relu_bwd up to the summation order (both were checked against the oracle at small sizes, tests/test_gpu_parity.py)."""
    pkg, ops, ctx = env
    n, c, h, w = 256, 64, 112, 112
    g = torch.Generator(device="cuda").manual_seed(5)
    X = torch.randn((n, h, w, c), device="cuda", generator=g) * 1.2 + 0.1
    scale = torch.rand((c,), device="cuda", generator=g) + 0.5
    scale[::7] *= -1.0     # negative scales: the max of the normalised values is not the max of x
    bias = 0.3 * torch.randn((c,), device="cuda", generator=g)
    rm1, rv1 = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    rm2, rv2 = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    yp, idx, sm, si = ops.batch_norm_relu_max_pool_forward_train(ctx, 0.9, X, scale, bias, rm1, rv1)
    y2, sm2, si2 = ops.batch_norm_2d_forward_train(ctx, 0.9, X, scale, bias, rm2, rv2, layout=pkg.ZB_NHWC, relu=True)
    yp2, idx2 = ops.max_pool_2d_indexed(ctx, y2, 3, 2, 1)
    for a, b in ((yp, yp2), (idx, idx2), (sm, sm2), (si, si2), (rm1, rm2), (rv1, rv2)):
        assert torch.equal(a, b)
    del yp2, idx2
    DYP = torch.randn(tuple(yp.shape), device="cuda", generator=g)
    dx, ds, db = ops.batch_norm_relu_max_pool_backward(ctx, X, DYP, idx, scale, bias, sm, si)
    dy2 = ops.max_pool_2d_indexed_backward(ctx, DYP, idx, tuple(y2.shape), 3, 2, 1)
    del y2
    dx2, ds2, db2 = ops.batch_norm_2d_relu_backward(ctx, X, dy2, scale, bias, sm, si, layout=pkg.ZB_NHWC)
    ctx.check()

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max())
    assert rel(dx, dx2) < 5e-5 and rel(ds, ds2) < 5e-5 and rel(db, db2) < 5e-5
