"""GPU parity tests proper: every call goes through the C ABI (libzenu_b200.so) and is compared with the CPU
oracle (oracle/) on the same seeded inputs, and with the reference's golden vectors (tests/golden/).

Tolerances (BASELINE.json north_star): rel 1e-3 for TF32 tensor-core math, 1e-5 for FFMA f32, 1e-10 for f64;
golden-vector tests use the reference tests' own max-abs tolerances.
"""
import ctypes
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import zenu_oracle as zo  # noqa: E402

pytestmark = pytest.mark.gpu

TOL = {"tf32": 1e-3, "tf32x3": 1e-5, "fp32": 1e-5, "f64": 1e-10}


@pytest.fixture(scope="module")
def zb():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from zenu_b200 import ops
    return ops


@pytest.fixture(scope="module")
def ctx(zb):
    c = zb.Context()
    yield c
    c.check()
    c.close()


@pytest.fixture(scope="module")
def lit(golden_dir):
    with open(os.path.join(golden_dir, "literals.json")) as f:
        return json.load(f)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def rel_err(got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    denom = np.linalg.norm(ref.ravel()) + 1e-300
    rel_l2 = np.linalg.norm((got - ref).ravel()) / denom
    max_rel = np.max(np.abs(got - ref)) / (np.max(np.abs(ref)) + 1e-300)
    return max(rel_l2, max_rel)


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64).ravel() - np.asarray(b, np.float64).ravel())))


def nhwc(a):
    return np.ascontiguousarray(np.transpose(a, (0, 2, 3, 1)))


def nchw(a):
    return np.ascontiguousarray(np.transpose(a, (0, 3, 1, 2)))


def math_of(zb, name):
    from zenu_b200 import ZB_MATH_FP32, ZB_MATH_TF32, ZB_MATH_TF32X3
    return {"tf32": ZB_MATH_TF32, "tf32x3": ZB_MATH_TF32X3, "fp32": ZB_MATH_FP32}[name]


# ------------------------------------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("math,tol", [("tf32x3", 1e-4), ("fp32", 1e-4), ("tf32", 5e-3)])
def test_conv_golden_json(zb, ctx, golden_dir, math, tol):
    # zenu-matrix/src/nn/conv/mod.rs:131-179: the reference's own tolerance (1e-4, max-abs) holds in the two f32-accurate math modes
    # (3xTF32 on the tensor cores, FFMA); plain TF32 -- since round 2 the C = 3 NCHW conv runs on the sliding-window tcgen05 kernels
    # instead of falling back to FFMA -- is held to its own rounding (10-bit mantissa operands, dw sums 1024 products of O(1) values)
    d = np.load(os.path.join(golden_dir, "conv2d.npz"))
    x, w = dev(d["input"]), dev(d["filter"])
    m = math_of(zb, math)
    y = zb.conv_fwd(ctx, x, w, pad=1, stride=1, dil=1, math=m)
    assert maxabs(host(y), d["output"]) < tol
    dy = torch.ones_like(y)
    assert maxabs(host(zb.conv_bkwd_data(ctx, dy, w, x.shape, 1, 1, 1, math=m)), d["grad_input"]) < tol
    scale = max(1.0, float(np.abs(d["grad_weight"]).max()))
    assert maxabs(host(zb.conv_bkwd_weight(ctx, dy, x, w.shape, 1, 1, 1, math=m)), d["grad_weight"]) < tol * scale


def test_conv_bias_golden_json(zb, ctx, golden_dir):
    from zenu_b200 import ZB_MATH_TF32X3
    d = np.load(os.path.join(golden_dir, "conv_bias.npz"))
    x, w, b = dev(d["input"]), dev(d["filter"]), dev(d["bias"])
    y = zb.conv_fwd(ctx, x, w, pad=1, stride=1, dil=1, bias=b, math=ZB_MATH_TF32X3)   # the reference's 1e-4 needs f32-level math
    assert maxabs(host(y), d["output"]) < 1e-4
    y2 = zb.conv2d_bias_add(ctx, zb.conv_fwd(ctx, x, w, pad=1, math=ZB_MATH_TF32X3), b)
    assert maxabs(host(y2), d["output"]) < 1e-4
    assert maxabs(host(zb.conv_fwd(ctx, x, w, pad=1, stride=1, dil=1, bias=b)), d["output"]) < 5e-3     # default math: TF32
    dy = torch.ones_like(y)
    assert maxabs(host(zb.conv2d_bias_bkwd(ctx, dy)), d["grad_bias"]) < 1e-4


def test_conv_small_literal(zb, ctx, lit):
    c = lit["conv_fwd_small"]
    x = dev(np.array(c["input"], np.float32).reshape(c["x_shape"]))
    w = dev(np.array(c["filter"], np.float32).reshape(c["w_shape"]))
    assert maxabs(host(zb.conv_fwd(ctx, x, w, pad=1)), c["output"]) < 1e-5


def test_bn_goldens(zb, ctx, lit):
    c = lit["bn_fwd_train"]
    x = dev(np.array(c["x"], np.float32).reshape(c["shape"]))
    rm, rv = dev(np.zeros(2, np.float32)), dev(np.zeros(2, np.float32))
    y, sm, si = zb.batch_norm_2d_forward_train(ctx, c["momentum"], x, dev(np.array(c["scale"], np.float32)),
                                               dev(np.array(c["bias"], np.float32)), rm, rv)
    for got, key in ((y, "y"), (rm, "running_mean"), (rv, "running_variance"), (sm, "saved_mean"), (si, "saved_inv_std")):
        assert maxabs(host(got), c[key]) < c["tol"], key
    c = lit["bn_bwd"]
    x = dev(np.array(c["x"], np.float32).reshape(c["shape"]))
    dy = dev(np.array(c["y_grad"], np.float32).reshape(c["shape"]))
    dx, ds, db = zb.batch_norm_2d_backward(ctx, x, dy, dev(np.array(c["scale"], np.float32)),
                                           dev(np.array(c["saved_mean"], np.float32)), dev(np.array(c["saved_inv_std"], np.float32)))
    assert maxabs(host(dx), c["x_grad"]) < c["tol"]
    assert maxabs(host(ds), c["scale_grad"]) < c["tol"]
    assert maxabs(host(db), c["bias_grad"]) < c["tol"]
    c = lit["bn_infer"]
    x = dev(np.array(c["x"], np.float32).reshape(c["shape"]))
    y = zb.batch_norm_2d_forward_inference(ctx, x, *[dev(np.array(c[k], np.float32)) for k in ("scale", "bias", "mean", "variance")])
    assert maxabs(host(y), c["y"]) < c["tol"]


def test_bn_autograd_golden(zb, ctx, lit):
    c = lit["bn_autograd"]
    x = dev(np.array(c["x"], np.float32).reshape(c["shape"]))
    dy = dev(np.array(c["y_grad"], np.float32).reshape(c["shape"]))
    scale, bias = dev(np.array(c["scale"], np.float32)), dev(np.array(c["bias"], np.float32))
    rm, rv = dev(np.array(c["prev_mean"], np.float32)), dev(np.array(c["prev_var"], np.float32))
    y, sm, si = zb.batch_norm_2d_forward_train(ctx, c["momentum"], x, scale, bias, rm, rv)
    assert maxabs(host(y), c["y"]) < c["tol_y"]
    dx, ds, db = zb.batch_norm_2d_backward(ctx, x, dy, scale, sm, si)
    assert maxabs(host(dx), c["x_grad"]) < c["tol_x_grad"]
    assert maxabs(host(ds), c["scale_grad"]) < c["tol_param_grad"]
    assert maxabs(host(db), c["bias_grad"]) < c["tol_param_grad"]
    dx2, _, _ = zb.batch_norm_2d_backward(ctx, x, dy, scale)  # saved stats None -> recomputed
    assert maxabs(host(dx2), c["x_grad"]) < c["tol_x_grad"]


def test_gemm_relu_literals(zb, ctx, lit):
    c = lit["gemm_3x4_4x5"]
    a, b = np.array(c["a"], np.float32).reshape(3, 4), np.array(c["b"], np.float32).reshape(4, 5)
    from zenu_b200 import ZB_MATH_FP32
    out = host(zb.gemm(ctx, dev(a), dev(b), math=ZB_MATH_FP32))
    assert float(np.abs(out.ravel() - np.array(c["c"])).sum()) < c["tol_asum"]
    c = lit["relu"]
    x = dev(np.array(c["x"], np.float32))
    assert maxabs(host(zb.relu(ctx, x)), c["y"]) < c["tol"]
    assert maxabs(host(zb.relu_backward_mask(ctx, x)), c["mask"]) < c["tol"]


# ------------------------------------------------------------------------------------------------ conv vs oracle
CONV_CASES = [
    # n, c, h, w, k, r, s, pad, stride, dil
    (2, 64, 14, 14, 128, 3, 3, 1, 1, 1),   # 3x3 s1 (im2col TMA)
    (2, 64, 15, 17, 64, 3, 3, 1, 2, 1),    # 3x3 s2, odd sizes
    (3, 128, 9, 9, 256, 1, 1, 0, 1, 1),    # 1x1 (plain GEMM)
    (2, 64, 16, 16, 96, 1, 1, 0, 2, 1),    # 1x1 s2 (downsample)
    (1, 32, 12, 12, 40, 3, 3, 2, 1, 2),    # dilation 2
    (2, 32, 8, 8, 32, 5, 5, 2, 1, 1),      # 5x5
    (2, 3, 20, 20, 16, 7, 7, 3, 2, 1),     # conv1-like: C=3 (FFMA path)
    (1, 20, 7, 7, 12, 3, 3, 0, 1, 1),      # ragged channels, no padding
    (3, 32, 9, 11, 192, 3, 3, 1, 1, 1),    # halo wgrad: two k tiles (the second half full), odd sizes, 3 images
    (2, 64, 10, 10, 40, 3, 3, 0, 1, 1),    # halo kernels without padding, ragged k tile
    (2, 32, 8, 12, 64, 1, 3, 1, 1, 1),     # 1x3 filter
    (2, 32, 12, 8, 32, 3, 1, 0, 1, 1),     # 3x1 filter
    (2, 3, 21, 19, 32, 5, 5, 2, 2, 1),     # small-C (stem) kernels: odd sizes, 5x5 s2
    (1, 4, 30, 40, 64, 3, 3, 1, 1, 1),     # small-C, C = 4, stride 1, two k chunks
    (2, 2, 17, 23, 32, 7, 7, 3, 3, 1),     # small-C, stride 3
    (1, 3, 18, 70, 128, 3, 3, 1, 2, 1),    # small-C, 128 output channels (4 dY boxes), three 32-pixel runs per row
    (2, 3, 12, 12, 96, 5, 5, 2, 1, 1),     # small-C, 96 output channels (ragged k tile), stride 1
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("math", ["tf32", "tf32x3", "fp32"])
def test_conv_vs_oracle(zb, ctx, case, layout, math):
    from zenu_b200 import ZB_NCHW, ZB_NHWC
    n, c, h, w, k, r, s, pad, stride, dil = case
    rng = np.random.default_rng(1234 + sum(case))
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    wt = (rng.standard_normal((k, c, r, s)) * np.sqrt(2.0 / (c * r * s))).astype(np.float32)
    y_ref = zo.conv2d_fwd(x.astype(np.float64), wt.astype(np.float64), pad, stride, dil)
    dy = rng.standard_normal(y_ref.shape).astype(np.float32)
    dx_ref = zo.conv2d_bkwd_data(dy.astype(np.float64), wt.astype(np.float64), x.shape, pad, stride, dil)
    dw_ref = zo.conv2d_bkwd_filter(dy.astype(np.float64), x.astype(np.float64), wt.shape, pad, stride, dil)
    m = math_of(zb, math)
    if layout == "nchw":
        L, X, W, DY, back_a, back_w = ZB_NCHW, dev(x), dev(wt), dev(dy), (lambda t: host(t)), (lambda t: host(t))
    else:
        L, X, W, DY = ZB_NHWC, dev(nhwc(x)), dev(nhwc(wt)), dev(nhwc(dy))
        back_a = back_w = lambda t: nchw(host(t))
    tol = TOL[math]
    y = zb.conv_fwd(ctx, X, W, pad, stride, dil, layout=L, math=m)
    assert rel_err(back_a(y), y_ref) < tol
    dx = zb.conv_bkwd_data(ctx, DY, W, X.shape, pad, stride, dil, layout=L, math=m)
    assert rel_err(back_a(dx), dx_ref) < tol
    dw = zb.conv_bkwd_weight(ctx, DY, X, W.shape, pad, stride, dil, layout=L, math=m)
    assert rel_err(back_w(dw), dw_ref) < tol
    ctx.check()


def _net_layers(arch, hw):
    from oracle import zenu_oracle_model as zm
    if arch == "small_cnn":
        return [("conv1", 3, 32, 3, 1, 1, hw), ("conv2", 32, 64, 3, 1, 1, hw)]
    blocks, _ = zm._resnet_plan(18 if arch == "resnet18" else 50)
    out = [("conv1", 3, 64, 7, 2, 3, hw)]
    h = (hw + 6 - 7) // 2 + 1
    h = (h + 2 - 3) // 2 + 1
    for name, convs, down in blocks:
        hin = h
        for i, (ci, co, k, stride, pad) in enumerate(convs):
            out.append((f"{name}.conv{i + 1}", ci, co, k, stride, pad, h))
            h = (h + 2 * pad - k) // stride + 1
        if down:
            out.append((f"{name}.down", down[0], down[1], 1, down[3], 0, hin))
    return out


@pytest.mark.parametrize("arch,n,hw", [("small_cnn", 64, 32), ("resnet18", 16, 64), ("resnet50", 16, 64)])
def test_conv_layers_vs_tf32_rounded_oracle(zb, ctx, arch, n, hw):
    """Every distinct conv geometry of the configs' networks (stem 7x7/s2 with C=3, 3x3 s1/s2, 1x1 s1/s2, up to 2048
    channels), fprop / dgrad / wgrad on the tcgen05 path, against the oracle fed with operands rounded to tf32
    (nearest-even, zo.tf32_round).  With the operand rounding modelled only the fp32 summation order differs, so the bound is
    5e-5 instead of the 1e-3 TF32 allowance: a kernel bug cannot hide inside the TF32 tolerance."""
    from zenu_b200 import ZB_MATH_TF32, ZB_NHWC
    zo.use_openblas()
    try:
        rng = np.random.default_rng(0)
        seen = set()
        for name, ci, co, k, stride, pad, h in _net_layers(arch, hw):
            key = (ci, co, k, stride, pad, h)
            if key in seen:
                continue
            seen.add(key)
            x = rng.standard_normal((n, ci, h, h)).astype(np.float32)
            w = (rng.standard_normal((co, ci, k, k)) * np.sqrt(2.0 / (ci * k * k))).astype(np.float32)
            xr, wr = zo.tf32_round(x, "rne"), zo.tf32_round(w, "rne")
            y_ref = zo.conv2d_fwd(xr, wr, pad, stride, 1)
            dy = rng.standard_normal(y_ref.shape).astype(np.float32)
            dyr = zo.tf32_round(dy, "rne")
            X, W, DY = dev(nhwc(x)), dev(nhwc(w)), dev(nhwc(dy))
            y = zb.conv_fwd(ctx, X, W, pad, stride, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
            assert rel_err(nchw(host(y)), y_ref) < 5e-5, (name, "fprop")
            dx = zb.conv_bkwd_data(ctx, DY, W, X.shape, pad, stride, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
            assert rel_err(nchw(host(dx)), zo.conv2d_bkwd_data(dyr, wr, x.shape, pad, stride, 1)) < 5e-5, (name, "dgrad")
            dw = zb.conv_bkwd_weight(ctx, DY, X, W.shape, pad, stride, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
            assert rel_err(nchw(host(dw)), zo.conv2d_bkwd_filter(dyr, xr, w.shape, pad, stride, 1)) < 5e-5, (name, "wgrad")
        ctx.check()
    finally:
        zo.use_plain_gemm()


@pytest.mark.parametrize("arch,n,hw", [("small_cnn", 16, 32), ("resnet50", 4, 64)])
def test_conv_layers_tf32x3(zb, ctx, arch, n, hw):
    """3xTF32 on every distinct conv geometry of the networks (incl. the C=3 stem's sliding-window path, the halo-reuse 3x3
    kernel, strided dgrad parity classes and split-K wgrad), against the oracle's exact arithmetic in f64: rel. 1e-5."""
    from zenu_b200 import ZB_MATH_TF32X3, ZB_NHWC
    rng = np.random.default_rng(5)
    seen = set()
    for name, ci, co, k, stride, pad, h in _net_layers(arch, hw):
        key = (ci, co, k, stride, pad, h)
        if key in seen:
            continue
        seen.add(key)
        x = rng.standard_normal((n, ci, h, h)).astype(np.float32)
        w = (rng.standard_normal((co, ci, k, k)) * np.sqrt(2.0 / (ci * k * k))).astype(np.float32)
        x64, w64 = x.astype(np.float64), w.astype(np.float64)
        y_ref = zo.conv2d_fwd(x64, w64, pad, stride, 1)
        dy = rng.standard_normal(y_ref.shape).astype(np.float32)
        X, W, DY = dev(nhwc(x)), dev(nhwc(w)), dev(nhwc(dy))
        y = zb.conv_fwd(ctx, X, W, pad, stride, 1, layout=ZB_NHWC, math=ZB_MATH_TF32X3)
        assert rel_err(nchw(host(y)), y_ref) < 1e-5, (name, "fprop")
        dx = zb.conv_bkwd_data(ctx, DY, W, X.shape, pad, stride, 1, layout=ZB_NHWC, math=ZB_MATH_TF32X3)
        assert rel_err(nchw(host(dx)), zo.conv2d_bkwd_data(dy.astype(np.float64), w64, x.shape, pad, stride, 1)) < 1e-5, (name, "dgrad")
        dw = zb.conv_bkwd_weight(ctx, DY, X, W.shape, pad, stride, 1, layout=ZB_NHWC, math=ZB_MATH_TF32X3)
        assert rel_err(nchw(host(dw)), zo.conv2d_bkwd_filter(dy.astype(np.float64), x64, w.shape, pad, stride, 1)) < 1e-5, (name, "wgrad")
        # accumulate form (residual fan-in): dx2 = base + dgrad
        if ci % 32 == 0 and co % 32 == 0:
            base = torch.ones_like(dx)
            zb.conv_bkwd_data_accumulate(ctx, DY, W, base, pad, stride, 1, layout=ZB_NHWC, math=ZB_MATH_TF32X3)
            assert rel_err(host(base) - 1.0, host(dx)) < 1e-4, (name, "dgrad accumulate")
            # masked fan-in: dx = dgrad + old (.) mask must equal accumulating onto the materialised product, bit for bit
            old = torch.randn_like(dx)
            words = torch.randint(-2 ** 31, 2 ** 31 - 1, ((old.numel() + 31) // 32,), dtype=torch.int64, device="cuda").to(torch.int32)
            want = zb.mask_apply(ctx, old, words)
            zb.conv_bkwd_data_accumulate(ctx, DY, W, want, pad, stride, 1, layout=ZB_NHWC, math=ZB_MATH_TF32X3)
            zb.conv_bkwd_data_accumulate_masked(ctx, DY, W, old, words, pad, stride, 1, layout=ZB_NHWC, math=ZB_MATH_TF32X3)
            np.testing.assert_array_equal(host(old), host(want), err_msg=f"{name} masked accumulate")
    ctx.check()


def test_conv_f64_vs_oracle(zb, ctx):
    n, c, h, w, k, r, s, pad, stride, dil = 2, 8, 10, 10, 6, 3, 3, 1, 2, 1
    rng = np.random.default_rng(7)
    x = rng.standard_normal((n, c, h, w))
    wt = rng.standard_normal((k, c, r, s))
    y_ref = zo.conv2d_fwd(x, wt, pad, stride, dil)
    dy = rng.standard_normal(y_ref.shape)
    y = zb.conv_fwd(ctx, dev(x), dev(wt), pad, stride, dil)
    assert rel_err(host(y), y_ref) < TOL["f64"]
    assert rel_err(host(zb.conv_bkwd_data(ctx, dev(dy), dev(wt), x.shape, pad, stride, dil)),
                   zo.conv2d_bkwd_data(dy, wt, x.shape, pad, stride, dil)) < TOL["f64"]
    assert rel_err(host(zb.conv_bkwd_weight(ctx, dev(dy), dev(x), wt.shape, pad, stride, dil)),
                   zo.conv2d_bkwd_filter(dy, x, wt.shape, pad, stride, dil)) < TOL["f64"]


def test_conv_bias_bwd_batched(zb, ctx):
    # N > 1: the reference GPU kernel is wrong here; the CPU semantics are the contract
    from zenu_b200 import ZB_NHWC
    rng = np.random.default_rng(3)
    dy = rng.standard_normal((5, 7, 6, 4)).astype(np.float32)
    ref = zo.conv2d_bias_bkwd(dy.astype(np.float64))
    assert rel_err(host(zb.conv2d_bias_bkwd(ctx, dev(dy))), ref) < 1e-5
    assert rel_err(host(zb.conv2d_bias_bkwd(ctx, dev(nhwc(dy)), layout=ZB_NHWC)), ref) < 1e-5


def test_conv_shape_errors(zb, ctx):
    from zenu_b200 import ZenuB200Error
    x = torch.zeros((1, 4, 8, 8), device="cuda")
    w = torch.zeros((2, 3, 3, 3), device="cuda")
    with pytest.raises(ZenuB200Error):
        zb.conv_fwd(ctx, x, w, pad=1)
    w = torch.zeros((2, 4, 11, 11), device="cuda")
    with pytest.raises(ZenuB200Error):
        zb.conv_fwd(ctx, x, w, pad=0)  # filter larger than input


# ------------------------------------------------------------------------------------------------ GEMM / Linear
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("math", ["tf32", "tf32x3", "fp32"])
@pytest.mark.parametrize("shape", [(256, 128, 64), (300, 200, 100), (64, 1000, 2048), (8, 8, 4096)])
def test_gemm_vs_oracle(zb, ctx, ta, tb, math, shape):
    m, n, k = shape
    rng = np.random.default_rng(m + n + k)
    a = rng.standard_normal((k, m) if ta else (m, k)).astype(np.float32)
    b = rng.standard_normal((n, k) if tb else (k, n)).astype(np.float32)
    c0 = rng.standard_normal((m, n)).astype(np.float32)
    ref = zo.gemm(a.astype(np.float64), b.astype(np.float64), ta, tb, 0.5, 0.25, c0.astype(np.float64))
    got = zb.gemm(ctx, dev(a), dev(b), ta, tb, 0.5, 0.25, dev(c0), math=math_of(zb, math))
    assert rel_err(host(got), ref) < TOL[math]
    ctx.check()


# CTA-pair (tcgen05 cta_group::2) launches: BN = 256 tiles, an even number of 128-row blocks and >= 16 K blocks per tile.  Ragged M / N / K,
# the ring wrapping several times, every operand-major combination (the MN-major forms are what wgrad and the pointwise dgrad use),
# alpha / beta through the deep old-tile variant, and the three math modes (3xTF32 adds the chained accumulator flushes).
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("math", ["tf32", "tf32x3"])
@pytest.mark.parametrize("shape", [(512, 256, 512), (500, 200, 520), (1024, 512, 2048), (256, 1000, 1024)])
def test_gemm_cta_pair_shapes(zb, ctx, ta, tb, math, shape):
    m, n, k = shape
    rng = np.random.default_rng(7 * m + 3 * n + k)
    a = rng.standard_normal((k, m) if ta else (m, k)).astype(np.float32)
    b = rng.standard_normal((n, k) if tb else (k, n)).astype(np.float32)
    c0 = rng.standard_normal((m, n)).astype(np.float32)
    ref = zo.gemm(a.astype(np.float64), b.astype(np.float64), ta, tb, 1.0, 0.0, c0.astype(np.float64))
    got = zb.gemm(ctx, dev(a), dev(b), ta, tb, 1.0, 0.0, None, math=math_of(zb, math))
    assert rel_err(host(got), ref) < TOL[math]
    ref = zo.gemm(a.astype(np.float64), b.astype(np.float64), ta, tb, 0.5, 0.25, c0.astype(np.float64))
    got = zb.gemm(ctx, dev(a), dev(b), ta, tb, 0.5, 0.25, dev(c0), math=math_of(zb, math))
    assert rel_err(host(got), ref) < TOL[math]
    ctx.check()


# convolutions that run on CTA pairs: pointwise 512 -> 256 (fprop: K-major filter; dgrad of 256 -> 512: MN-major filter),
# a strided 3x3 through im2col loads (fprop, the dgrad parity classes with enough taps), wgrad with MN-major im2col loads
@pytest.mark.parametrize("case", [(2, 512, 14, 14, 256, 1, 0, 1), (2, 256, 14, 14, 512, 1, 0, 1), (4, 256, 15, 17, 256, 3, 1, 2),
                                  (3, 512, 9, 9, 512, 1, 0, 2)])
def test_conv_cta_pair_shapes(zb, ctx, case):
    from zenu_b200 import ZB_NHWC
    n, c, h, w, k, r, pad, stride = case
    rng = np.random.default_rng(99 + sum(case))
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    wt = (rng.standard_normal((k, c, r, r)) * np.sqrt(2.0 / (c * r * r))).astype(np.float32)
    y_ref = zo.conv2d_fwd(x.astype(np.float64), wt.astype(np.float64), pad, stride, 1)
    dy = rng.standard_normal(y_ref.shape).astype(np.float32)
    dx_ref = zo.conv2d_bkwd_data(dy.astype(np.float64), wt.astype(np.float64), x.shape, pad, stride, 1)
    dw_ref = zo.conv2d_bkwd_filter(dy.astype(np.float64), x.astype(np.float64), wt.shape, pad, stride, 1)
    X, W, DY = dev(nhwc(x)), dev(nhwc(wt)), dev(nhwc(dy))
    m = math_of(zb, "tf32")
    y = zb.conv_fwd(ctx, X, W, pad, stride, 1, layout=ZB_NHWC, math=m)
    assert rel_err(nchw(host(y)), y_ref) < TOL["tf32"]
    dx = zb.conv_bkwd_data(ctx, DY, W, X.shape, pad, stride, 1, layout=ZB_NHWC, math=m)
    assert rel_err(nchw(host(dx)), dx_ref) < TOL["tf32"]
    dw = zb.conv_bkwd_weight(ctx, DY, X, W.shape, pad, stride, 1, layout=ZB_NHWC, math=m)
    assert rel_err(nchw(host(dw)), dw_ref) < TOL["tf32"]
    ctx.check()


@pytest.mark.parametrize("math", ["tf32", "tf32x3", "fp32"])
def test_linear_vs_oracle(zb, ctx, math):
    rng = np.random.default_rng(11)
    b, i, o = 64, 512, 10
    x = rng.standard_normal((b, i)).astype(np.float32)
    w = (rng.standard_normal((o, i)) / np.sqrt(i)).astype(np.float32)
    bias = rng.standard_normal((o,)).astype(np.float32)
    dy = rng.standard_normal((b, o)).astype(np.float32)
    m = math_of(zb, math)
    y = zb.linear_fwd(ctx, dev(x), dev(w), dev(bias), math=m)
    assert rel_err(host(y), zo.linear_fwd(x.astype(np.float64), w.astype(np.float64), bias.astype(np.float64))) < TOL[math]
    dx, dw, db = zb.linear_bwd(ctx, dev(x), dev(w), dev(dy), math=m)
    rdx, rdw, rdb = zo.linear_bwd(x.astype(np.float64), w.astype(np.float64), dy.astype(np.float64))
    assert rel_err(host(dx), rdx) < TOL[math]
    assert rel_err(host(dw), rdw) < TOL[math]
    assert rel_err(host(db), rdb) < 1e-5


# ------------------------------------------------------------------------------------------------ BatchNorm
@pytest.mark.parametrize("shape", [(4, 64, 9, 7), (3, 5, 6, 6), (2, 256, 4, 4), (16, 8, 1, 1)])
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_bn_vs_oracle(zb, ctx, shape, layout, dtype):
    from zenu_b200 import ZB_NCHW, ZB_NHWC
    rng = np.random.default_rng(sum(shape))
    n, c, h, w = shape
    x = (rng.standard_normal(shape) * 2.0 + 3.0).astype(dtype)   # non-zero mean exercises the shifted statistics
    dy = rng.standard_normal(shape).astype(dtype)
    scale, bias = rng.standard_normal(c).astype(dtype), rng.standard_normal(c).astype(dtype)
    rm0, rv0 = rng.standard_normal(c).astype(dtype), (rng.random(c) + 0.5).astype(dtype)
    x64, dy64 = x.astype(np.float64), dy.astype(np.float64)
    y_ref, rm_ref, rv_ref, sm_ref, si_ref = zo.bn2d_fwd_train(x64, scale.astype(np.float64), bias.astype(np.float64),
                                                                rm0.astype(np.float64), rv0.astype(np.float64), 0.9)
    dx_ref, ds_ref, db_ref = zo.bn2d_bwd(x64, dy64, scale.astype(np.float64), sm_ref, si_ref)
    tol = 2e-5 if dtype == np.float32 else 1e-10
    if layout == "nchw":
        L, X, DY, back = ZB_NCHW, dev(x), dev(dy), host
    else:
        L, X, DY, back = ZB_NHWC, dev(nhwc(x)), dev(nhwc(dy)), (lambda t: nchw(host(t)))
    rm, rv = dev(rm0), dev(rv0)
    y, sm, si = zb.batch_norm_2d_forward_train(ctx, 0.9, X, dev(scale), dev(bias), rm, rv, layout=L)
    assert rel_err(back(y), y_ref) < tol
    assert rel_err(host(rm), rm_ref) < tol and rel_err(host(rv), rv_ref) < tol
    assert rel_err(host(sm), sm_ref) < tol and rel_err(host(si), si_ref) < tol
    dx, ds, db = zb.batch_norm_2d_backward(ctx, X, DY, dev(scale), sm, si, layout=L)
    assert rel_err(back(dx), dx_ref) < 20 * tol
    assert rel_err(host(ds), ds_ref) < 20 * tol and rel_err(host(db), db_ref) < 20 * tol
    yi = zb.batch_norm_2d_forward_inference(ctx, X, dev(scale), dev(bias), dev(rm0), dev(rv0), layout=L)
    assert rel_err(back(yi), zo.bn2d_fwd_infer(x64, scale.astype(np.float64), bias.astype(np.float64),
                                                 rm0.astype(np.float64), rv0.astype(np.float64))) < tol


@pytest.mark.parametrize("case", [
    (4, 64, 14, 14, 256, 1, 0, 1),     # 1x1 GEMM path, 2 column tiles of 128 -> one of 256
    (3, 64, 18, 18, 64, 3, 1, 1),      # halo-reuse 3x3 (dead halo rows must not enter the statistics)
    (2, 128, 15, 15, 512, 1, 0, 2),    # strided 1x1 (im2col path), ragged last M tile
    (2, 32, 16, 16, 96, 3, 1, 2),      # 3x3 stride 2 (im2col path), N = 96 (BN 128 tile, partly empty)
    (2, 3, 40, 40, 64, 7, 3, 2),       # C = 3 stem (sliding-window path)
])
def test_conv_fused_bn_statistics(zb, ctx, case):
    """conv fprop + BatchNorm statistics fused in the conv epilogue, then BN apply from those partials == the oracle's
    conv -> BatchNorm(train) on the same inputs; also == the unfused GPU path (same y, same saved statistics)."""
    from zenu_b200 import ZB_NHWC
    n, c, hw, _, k, r, pad, stride = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((n, c, hw, hw)).astype(np.float32)
    w = (rng.standard_normal((k, c, r, r)) * np.sqrt(2.0 / (c * r * r))).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, k).astype(np.float32)
    bias = rng.standard_normal(k).astype(np.float32)
    rmean = (0.3 * rng.standard_normal(k)).astype(np.float32)      # non-trivial shift
    rvar = rng.uniform(0.5, 2.0, k).astype(np.float32)
    X, W = dev(nhwc(x)), dev(nhwc(w))
    rm1, rv1, rm2, rv2 = dev(rmean), dev(rvar), dev(rmean), dev(rvar)
    y, partial, rows = zb.conv_fwd_bnstats(ctx, X, W, rm1, pad, stride, 1, layout=ZB_NHWC)
    assert rows > 0, "statistics were not fused for this shape"
    a, sm, si = zb.batch_norm_2d_forward_train_prestats(ctx, 0.9, y, dev(scale), dev(bias), rm1, rv1, partial, rows, rm1, relu=True)
    y2 = zb.conv_fwd(ctx, X, W, pad, stride, 1, layout=ZB_NHWC)
    a2, sm2, si2 = zb.batch_norm_2d_forward_train(ctx, 0.9, y2, dev(scale), dev(bias), rm2, rv2, layout=ZB_NHWC, relu=True)
    np.testing.assert_array_equal(host(y), host(y2))
    assert rel_err(host(sm), host(sm2)) < 1e-5 and rel_err(host(si), host(si2)) < 1e-5
    assert rel_err(host(rm1), host(rm2)) < 1e-5 and rel_err(host(rv1), host(rv2)) < 1e-5
    assert rel_err(host(a), host(a2)) < 1e-5
    # against the oracle BatchNorm fed with the SAME conv output (conv parity itself is covered above)
    ref, rm_ref, rv_ref, sm_ref, si_ref = zo.bn2d_fwd_train(nchw(host(y)), scale, bias, rmean.copy(), rvar.copy(), 0.9)
    assert rel_err(nchw(host(a)), zo.relu(ref)) < 1e-4
    assert rel_err(host(sm), sm_ref) < 1e-4 and rel_err(host(si), si_ref) < 1e-4
    assert rel_err(host(rm1), rm_ref) < 1e-4 and rel_err(host(rv1), rv_ref) < 1e-4
    ctx.check()


@pytest.mark.parametrize("case", [
    (2, 64, 32, 32, 64, 3, 1),       # 3x3 / 2 / pad 1: Q = 16, three row blocks per image
    (3, 32, 33, 31, 128, 3, 1),      # odd extents (last window row / column partly in the padding), 128 output channels
    (2, 64, 36, 36, 96, 5, 2),       # 5x5 / 2 / pad 2: plane offsets 0..2, N = 96 (partly empty BN = 128 tile)
    (2, 128, 30, 34, 256, 3, 0),     # no padding: both parity planes start at 0 / 1, streamed filter, CTA pairs
])
def test_conv_stride2_halo_planes(zb, ctx, case):
    """Stride-2 RxS forward convs on the halo-reuse kernel: the input's four (row, column) parity planes land as dense rasters through ONE
    tensor map with element strides (1, 2, 2, 1), tap (r, s) = a raster offset inside plane ((r - pad) mod 2, (s - pad) mod 2).  The plan
    must say so (planes=4), the result must match the tf32-rounded oracle (5e-5), 3xTF32 exact arithmetic (1e-5), the fused BatchNorm
    statistics the values stored, and the im2col form (ZENU_B200_NO_HALO_S2 is the A/B switch; here: the FFMA kernel as second opinion)."""
    from zenu_b200 import ZB_MATH_FP32, ZB_MATH_TF32, ZB_MATH_TF32X3, ZB_NHWC
    n, c, h, w_, k, r, pad = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((n, c, h, w_)).astype(np.float32)
    wt = (rng.standard_normal((k, c, r, r)) * np.sqrt(2.0 / (c * r * r))).astype(np.float32)
    X, W = dev(nhwc(x)), dev(nhwc(wt))
    plan = zb.conv_plan_describe(ctx, zb.PLAN_FPROP, tuple(X.shape), tuple(W.shape), pad, 2, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
    assert "halo_conv" in plan and "planes=4" in plan, plan
    shift = dev((0.1 * rng.standard_normal(k)).astype(np.float32))
    y, partial, rows = zb.conv_fwd_bnstats(ctx, X, W, shift, pad, 2, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
    assert rows > 0
    xr, wr = zo.tf32_round(x, "rne"), zo.tf32_round(wt, "rne")
    y_ref = zo.conv2d_fwd(xr, wr, pad, 2, 1)
    y_h = nchw(host(y))
    assert rel_err(y_h, y_ref) < 5e-5
    part = host(partial)[:rows].astype(np.float64).sum(axis=0)
    dlt = y_h.astype(np.float64) - host(shift).astype(np.float64)[None, :, None, None]
    np.testing.assert_allclose(part[0], dlt.sum(axis=(0, 2, 3)), rtol=1e-4, atol=1e-3 * np.sqrt(dlt[:, 0].size))
    np.testing.assert_allclose(part[1], (dlt * dlt).sum(axis=(0, 2, 3)), rtol=1e-4)
    y3 = zb.conv_fwd(ctx, X, W, pad, 2, 1, layout=ZB_NHWC, math=ZB_MATH_TF32X3)
    assert rel_err(nchw(host(y3)), zo.conv2d_fwd(x.astype(np.float64), wt.astype(np.float64), pad, 2, 1)) < 1e-5
    yf = zb.conv_fwd(ctx, X, W, pad, 2, 1, layout=ZB_NHWC, math=ZB_MATH_FP32)
    assert rel_err(host(y3), host(yf)) < 1e-5
    bias = dev(rng.standard_normal(k).astype(np.float32))
    yb = zb.conv_fwd(ctx, X, W, pad, 2, 1, bias=bias, layout=ZB_NHWC, math=ZB_MATH_TF32)
    assert rel_err(host(yb), host(y) + host(bias)[None, None, None, :]) < 1e-6
    # backward data: every parity class with two or more taps is a stride-1 conv over dY on the same kernel (custom tap set, outputs
    # scattered to the class' pixels); plain, accumulating and masked-accumulating forms
    dy = rng.standard_normal(y_ref.shape).astype(np.float32)
    DY = dev(nhwc(dy))
    dplan = zb.conv_plan_describe(ctx, zb.PLAN_DGRAD, tuple(X.shape), tuple(W.shape), pad, 2, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
    assert "halo_conv" in dplan, dplan
    dx = zb.conv_bkwd_data(ctx, DY, W, X.shape, pad, 2, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
    assert rel_err(nchw(host(dx)), zo.conv2d_bkwd_data(zo.tf32_round(dy, "rne"), wr, x.shape, pad, 2, 1)) < 5e-5
    dx3 = zb.conv_bkwd_data(ctx, DY, W, X.shape, pad, 2, 1, layout=ZB_NHWC, math=ZB_MATH_TF32X3)
    assert rel_err(nchw(host(dx3)), zo.conv2d_bkwd_data(dy.astype(np.float64), wt.astype(np.float64), x.shape, pad, 2, 1)) < 1e-5
    old = torch.randn_like(dx)
    acc = old.clone()
    zb.conv_bkwd_data_accumulate(ctx, DY, W, acc, pad, 2, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
    assert rel_err(host(acc) - host(old), host(dx)) < 1e-4
    words = torch.randint(-2 ** 31, 2 ** 31 - 1, ((old.numel() + 31) // 32,), dtype=torch.int64, device="cuda").to(torch.int32)
    want = zb.mask_apply(ctx, old, words)
    zb.conv_bkwd_data_accumulate(ctx, DY, W, want, pad, 2, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
    zb.conv_bkwd_data_accumulate_masked(ctx, DY, W, old, words, pad, 2, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
    np.testing.assert_array_equal(host(old), host(want))
    ctx.check()


@pytest.mark.parametrize("case", [
    (5, 64, 7, 7, 64, 3, 1),        # 7x7, resident filter, odd batch: the last tile holds ONE image (the other is TMA zero fill)
    (6, 128, 7, 7, 512, 3, 1),      # streamed filter, two column blocks, CTA pairs
    (9, 64, 3, 7, 64, 3, 1),        # 3x7: four images per tile, the last tile holds one
    (7, 32, 7, 3, 96, 3, 1),        # 7x3: four images per tile, N = 96 (partly empty BN = 128 column block)
])
def test_conv_small_maps_stacked_tiles(zb, ctx, case):
    """Small feature maps (the 7x7 stage) on the halo kernel with several images per 128-row tile: the rasters are stacked with the
    padding rows / columns shared between neighbouring rows and images.  fprop (+ bias, + fused statistics) and dgrad (plain, accumulate,
    masked accumulate) against the tf32-rounded oracle (5e-5) and 3xTF32 against exact arithmetic (1e-5); the plan must name the mode."""
    from zenu_b200 import ZB_MATH_TF32, ZB_MATH_TF32X3, ZB_NHWC
    n, c, h, w_, k, r, pad = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((n, c, h, w_)).astype(np.float32)
    wt = (rng.standard_normal((k, c, r, r)) * np.sqrt(2.0 / (c * r * r))).astype(np.float32)
    X, W = dev(nhwc(x)), dev(nhwc(wt))
    plan = zb.conv_plan_describe(ctx, zb.PLAN_FPROP, tuple(X.shape), tuple(W.shape), pad, 1, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
    assert "halo_conv" in plan and "stack=1" not in plan, plan
    shift = dev((0.1 * rng.standard_normal(k)).astype(np.float32))
    y, partial, rows = zb.conv_fwd_bnstats(ctx, X, W, shift, pad, 1, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
    assert rows > 0
    xr, wr = zo.tf32_round(x, "rne"), zo.tf32_round(wt, "rne")
    y_ref = zo.conv2d_fwd(xr, wr, pad, 1, 1)
    y_h = nchw(host(y))
    assert rel_err(y_h, y_ref) < 5e-5
    part = host(partial)[:rows].astype(np.float64).sum(axis=0)
    dlt = y_h.astype(np.float64) - host(shift).astype(np.float64)[None, :, None, None]
    np.testing.assert_allclose(part[0], dlt.sum(axis=(0, 2, 3)), rtol=1e-4, atol=1e-3 * np.sqrt(dlt[:, 0].size))
    np.testing.assert_allclose(part[1], (dlt * dlt).sum(axis=(0, 2, 3)), rtol=1e-4)
    y3 = zb.conv_fwd(ctx, X, W, pad, 1, 1, layout=ZB_NHWC, math=ZB_MATH_TF32X3)
    assert rel_err(nchw(host(y3)), zo.conv2d_fwd(x.astype(np.float64), wt.astype(np.float64), pad, 1, 1)) < 1e-5
    dy = rng.standard_normal(y_ref.shape).astype(np.float32)
    DY = dev(nhwc(dy))
    dplan = zb.conv_plan_describe(ctx, zb.PLAN_DGRAD, tuple(X.shape), tuple(W.shape), pad, 1, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
    assert "halo_conv" in dplan and "stack=1" not in dplan, dplan
    dx = zb.conv_bkwd_data(ctx, DY, W, X.shape, pad, 1, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
    assert rel_err(nchw(host(dx)), zo.conv2d_bkwd_data(zo.tf32_round(dy, "rne"), wr, x.shape, pad, 1, 1)) < 5e-5
    if c % 32 == 0:
        old = torch.randn_like(dx)
        words = torch.randint(-2 ** 31, 2 ** 31 - 1, ((old.numel() + 31) // 32,), dtype=torch.int64, device="cuda").to(torch.int32)
        want = zb.mask_apply(ctx, old, words)
        zb.conv_bkwd_data_accumulate(ctx, DY, W, want, pad, 1, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
        assert rel_err(host(want) - host(zb.mask_apply(ctx, old, words)), host(dx)) < 1e-4
        zb.conv_bkwd_data_accumulate_masked(ctx, DY, W, old, words, pad, 1, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
        np.testing.assert_array_equal(host(old), host(want))
    # backward filter: the pixel reduction of the halo wgrad kernel walks the same stacked rasters (dY's rows >= P and columns >= Q, and
    # the images past the end of the batch, are TMA zero fill)
    wplan = zb.conv_plan_describe(ctx, zb.PLAN_WGRAD, tuple(X.shape), tuple(W.shape), pad, 1, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
    assert "wgrad_halo" in wplan and "stack=1 " not in wplan, wplan
    dw = zb.conv_bkwd_weight(ctx, DY, X, W.shape, pad, 1, 1, layout=ZB_NHWC, math=ZB_MATH_TF32)
    assert rel_err(nchw(host(dw)), zo.conv2d_bkwd_filter(zo.tf32_round(dy, "rne"), xr, wt.shape, pad, 1, 1)) < 5e-5
    dw3 = zb.conv_bkwd_weight(ctx, DY, X, W.shape, pad, 1, 1, layout=ZB_NHWC, math=ZB_MATH_TF32X3)
    assert rel_err(nchw(host(dw3)), zo.conv2d_bkwd_filter(dy.astype(np.float64), x.astype(np.float64), wt.shape, pad, 1, 1)) < 1e-5
    ctx.check()


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_bn_fused_relu_residual(zb, ctx, layout):
    """Fused BN+add+ReLU fwd/bwd == the reference's separate nodes (BN -> add -> relu) composed from oracle ops."""
    from zenu_b200 import ZB_NCHW, ZB_NHWC
    rng = np.random.default_rng(99)
    shape = (4, 32, 6, 5)
    c = shape[1]
    x = rng.standard_normal(shape)
    res = rng.standard_normal(shape)
    dy = rng.standard_normal(shape)
    scale, bias = rng.standard_normal(c), rng.standard_normal(c)
    bn, _, _, sm, si = zo.bn2d_fwd_train(x, scale, bias, np.zeros(c), np.ones(c), 0.9)
    z = zo.ewise("add", bn, res)
    out_ref = zo.relu(z)
    dz = zo.ewise("mul", dy, zo.relu_backward_mask(z))          # relu backward (activation/relu.rs:36-46)
    dx_ref, ds_ref, db_ref = zo.bn2d_bwd(x, dz, scale, sm, si)  # add backward passes dz to both inputs
    f = np.float32
    if layout == "nchw":
        L, cv, back = ZB_NCHW, (lambda a: dev(a.astype(f))), host
    else:
        L, cv, back = ZB_NHWC, (lambda a: dev(nhwc(a.astype(f)))), (lambda t: nchw(host(t)))
    X, R, DY = cv(x), cv(res), cv(dy)
    S, B = dev(scale.astype(f)), dev(bias.astype(f))
    y, smg, sig = zb.batch_norm_2d_forward_train(ctx, 0.9, X, S, B, dev(np.zeros(c, f)), dev(np.ones(c, f)), layout=L,
                                                 residual=R, relu=True)
    assert rel_err(back(y), out_ref) < 2e-5
    dx, ds, db, dres = zb.batch_norm_2d_backward(ctx, X, DY, S, smg, sig, layout=L, y=y, want_residual_grad=True)
    assert rel_err(back(dx), dx_ref) < 2e-4
    assert rel_err(back(dres), dz) < 2e-5
    assert rel_err(host(ds), ds_ref) < 2e-4 and rel_err(host(db), db_ref) < 2e-4


@pytest.mark.parametrize("shape", [(4, 64, 9, 7), (3, 32, 5, 5), (2, 256, 6, 6)])
def test_bn_add_relu_bit_mask(zb, ctx, shape):
    """Fused BN+add+ReLU whose forward writes a 1-bit ReLU mask and whose backward reads it instead of y: identical results to the
    y-based fused path (bit-exact) and equal to the reference's separate batch_norm -> add -> relu nodes (oracle)."""
    rng = np.random.default_rng(sum(shape))
    n, c, h, w = shape
    x = rng.standard_normal(shape).astype(np.float32)
    res = rng.standard_normal(shape).astype(np.float32)
    dy = rng.standard_normal(shape).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, c).astype(np.float32)
    bias = (0.2 * rng.standard_normal(c)).astype(np.float32)
    from zenu_b200 import ZB_NHWC
    X, R, DY = dev(nhwc(x)), dev(nhwc(res)), dev(nhwc(dy))
    rm, rv = dev(np.zeros(c, np.float32)), dev(np.ones(c, np.float32))
    y, sm, si, mask = zb.batch_norm_2d_forward_train_masked(ctx, 0.9, X, dev(scale), dev(bias), rm, rv, residual=R)
    rm2, rv2 = dev(np.zeros(c, np.float32)), dev(np.ones(c, np.float32))
    y2, sm2, si2 = zb.batch_norm_2d_forward_train(ctx, 0.9, X, dev(scale), dev(bias), rm2, rv2, layout=ZB_NHWC, residual=R, relu=True)
    np.testing.assert_array_equal(host(y), host(y2))
    bits = np.unpackbits(host(mask).view(np.uint8), bitorder="little")[: x.size]
    np.testing.assert_array_equal(bits.astype(bool), (host(y).ravel() > 0))
    dx, ds, db, dres = zb.batch_norm_2d_backward_masked(ctx, X, DY, dev(scale), sm, si, mask)
    dx2, ds2, db2, dres2 = zb.batch_norm_2d_backward(ctx, X, DY, dev(scale), sm2, si2, layout=ZB_NHWC, y=y2, want_residual_grad=True)
    for a, b in ((dx, dx2), (ds, ds2), (db, db2), (dres, dres2)):
        np.testing.assert_array_equal(host(a), host(b))
    # oracle: separate nodes
    bn_ref, _, _, sm_ref, si_ref = zo.bn2d_fwd_train(x, scale, bias, np.zeros(c, np.float32), np.ones(c, np.float32), 0.9)
    out_ref = zo.relu(zo.ewise("add", bn_ref, res))
    assert rel_err(nchw(host(y)), out_ref) < 1e-5
    g = zo.ewise("mul", dy, (out_ref > 0).astype(np.float32))
    dx_ref, ds_ref, db_ref = zo.bn2d_bwd(x, g, scale, sm_ref, si_ref)
    assert rel_err(nchw(host(dx)), dx_ref) < 1e-4 and rel_err(host(ds), ds_ref) < 1e-4 and rel_err(host(db), db_ref) < 1e-4
    assert rel_err(nchw(host(dres)), g) < 1e-6
    # lazy form: the masked gradient is not written out, everything else is bit-identical; a plain BN (downsample branch) reading
    # (dy, mask) lazily equals the same BN fed the materialised product
    dx3, ds3, db3, none = zb.batch_norm_2d_backward_masked(ctx, X, DY, dev(scale), sm, si, mask, want_residual_grad=False)
    assert none is None
    for a, b in ((dx3, dx), (ds3, ds), (db3, db)):
        np.testing.assert_array_equal(host(a), host(b))
    np.testing.assert_array_equal(host(zb.mask_apply(ctx, DY, mask)), host(dres))
    dx4, ds4, db4 = zb.batch_norm_2d_backward(ctx, X, dres, dev(scale), sm, si, layout=ZB_NHWC)
    assert rel_err(host(dx3), host(dx4)) < 1e-5 and rel_err(host(ds3), host(ds4)) < 1e-5 and rel_err(host(db3), host(db4)) < 1e-5
    ctx.check()


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_bn_relu_backward_recomputed_mask(zb, ctx, layout, dtype):
    """zb_bn2d_relu_bwd (mask recomputed from x) == BN -> relu backward composed from oracle ops, and is bit-identical
    to the y-reading variant."""
    from zenu_b200 import ZB_NCHW, ZB_NHWC
    rng = np.random.default_rng(314)
    shape = (6, 24, 7, 9)
    c = shape[1]
    f = np.float32 if dtype == "f32" else np.float64
    x = rng.standard_normal(shape).astype(f)
    dy = rng.standard_normal(shape).astype(f)
    scale, bias = rng.standard_normal(c).astype(f), (0.3 * rng.standard_normal(c)).astype(f)
    bn, _, _, sm, si = zo.bn2d_fwd_train(x.astype(np.float64), scale.astype(np.float64), bias.astype(np.float64), np.zeros(c), np.ones(c), 0.9)
    dz = zo.ewise("mul", dy.astype(np.float64), zo.relu_backward_mask(bn))
    dx_ref, ds_ref, db_ref = zo.bn2d_bwd(x.astype(np.float64), dz, scale.astype(np.float64), sm, si)
    if layout == "nchw":
        L, cv, back = ZB_NCHW, dev, host
    else:
        L, cv, back = ZB_NHWC, (lambda a: dev(nhwc(a))), (lambda t: nchw(host(t)))
    X, DY, S, B = cv(x), cv(dy), dev(scale), dev(bias)
    y, smg, sig = zb.batch_norm_2d_forward_train(ctx, 0.9, X, S, B, dev(np.zeros(c, f)), dev(np.ones(c, f)), layout=L, relu=True)
    dx, ds, db = zb.batch_norm_2d_relu_backward(ctx, X, DY, S, B, smg, sig, layout=L)
    tol = 2e-4 if dtype == "f32" else 1e-10
    # elements whose pre-activation is within rounding of zero may flip against the float64 oracle: compare in L2
    assert rel_err(back(dx), dx_ref) < tol and rel_err(host(ds), ds_ref) < tol and rel_err(host(db), db_ref) < tol
    dx2, ds2, db2 = zb.batch_norm_2d_backward(ctx, X, DY, S, smg, sig, layout=L, y=y)
    np.testing.assert_array_equal(host(dx), host(dx2))
    np.testing.assert_array_equal(host(ds), host(ds2))
    np.testing.assert_array_equal(host(db), host(db2))


def test_dgrad_accumulate(zb, ctx):
    """dx += dgrad (fan-in folded into the epilogue) == dgrad + old, stride 1 and stride 2 (scatter epilogue)."""
    from zenu_b200 import ZB_NHWC
    rng = np.random.default_rng(77)
    for (n, c, h, w, k, r, pad, stride) in ((2, 64, 12, 12, 64, 3, 1, 1), (2, 64, 13, 13, 32, 1, 0, 2), (2, 32, 9, 9, 64, 3, 1, 2)):
        x_shape = (n, h, w, c)
        wt = (rng.standard_normal((k, r, r, c)) * 0.1).astype(np.float32)
        p = (h + 2 * pad - r) // stride + 1
        dy = rng.standard_normal((n, p, p, k)).astype(np.float32)
        old = rng.standard_normal(x_shape).astype(np.float32)
        base = host(zb.conv_bkwd_data(ctx, dev(dy), dev(wt), x_shape, pad, stride, 1, layout=ZB_NHWC))
        acc = dev(old.copy())
        zb.conv_bkwd_data_accumulate(ctx, dev(dy), dev(wt), acc, pad, stride, 1, layout=ZB_NHWC)
        np.testing.assert_allclose(host(acc), base + old, rtol=1e-6, atol=1e-6)
    ctx.check()


def test_maxpool_indexed(zb, ctx):
    """Indexed NHWC max-pool: forward == oracle (zero padding takes part, first max wins), backward gather == oracle."""
    rng = np.random.default_rng(18)
    for shape, (k, s, p) in (((2, 8, 11, 9), (3, 2, 1)), ((1, 4, 8, 8), (2, 2, 0)), ((2, 12, 7, 10), (3, 1, 1)), ((1, 4, 9, 12), (5, 2, 2))):
        x = np.maximum(rng.standard_normal(shape), 0).astype(np.float32) - (0.2 if k == 3 and s == 1 else 0.0)
        x = x.astype(np.float32)
        y_ref = zo.maxpool2d_fwd(x, k, s, p)
        dy = rng.standard_normal(y_ref.shape).astype(np.float32)
        dx_ref = zo.maxpool2d_bwd(x, dy, k, s, p)
        y, idx = zb.max_pool_2d_indexed(ctx, dev(nhwc(x)), k, s, p)
        np.testing.assert_array_equal(nchw(host(y)), y_ref)
        dx = zb.max_pool_2d_indexed_backward(ctx, dev(nhwc(dy)), idx, nhwc(x).shape, k, s, p)
        np.testing.assert_allclose(nchw(host(dx)), dx_ref, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("shape,prestats", [((3, 64, 12, 10), False), ((2, 16, 9, 11), False), ((4, 64, 16, 16), True), ((2, 256, 6, 8), False)])
def test_bn_relu_maxpool_fused_stem(zb, ctx, shape, prestats):
    """zb_bn2d_relu_maxpool_fwd_train / _bwd (the ResNet stem as one pass each way) against (a) the separate entry points they replace
    - forward bit for bit (pooled values, winner codes, saved and running statistics), backward to rounding (other summation order) -
    and (b) the oracle's batch_norm -> relu -> max_pool nodes (zenu-matrix/src/nn/batch_norm.rs:283-420, nn/pool2d semantics:
    zero padding takes part, first max wins).  Odd and even extents; statistics reduced here or handed over by the conv epilogue."""
    from zenu_b200 import ZB_MATH_TF32, ZB_NHWC
    rng = np.random.default_rng(sum(shape))
    n, c, h, w = shape
    x = (rng.standard_normal(shape) * 1.3 + 0.2).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, c).astype(np.float32) * np.where(rng.random(c) < 0.2, -1.0, 1.0).astype(np.float32)
    bias = (0.3 * rng.standard_normal(c)).astype(np.float32)
    X, S, B = dev(nhwc(x)), dev(scale), dev(bias)
    kw = {}
    if prestats:   # statistics from a conv epilogue: a 1x1 identity conv reproduces x and hands its partial sums over
        eye = dev(np.eye(c, dtype=np.float32).reshape(c, 1, 1, c))
        shift = dev((0.05 * rng.standard_normal(c)).astype(np.float32))
        xs, partial, rows = zb.conv_fwd_bnstats(ctx, X, eye, shift, pad=0, stride=1, dil=1, layout=ZB_NHWC, math=ZB_MATH_TF32)
        assert rows > 0
        X = xs
        x = nchw(host(X))
        kw = dict(stat_partial=partial, stat_rows=rows, shift=shift)
    rm1, rv1 = dev(np.zeros(c, np.float32)), dev(np.ones(c, np.float32))
    yp, idx, sm, si = zb.batch_norm_relu_max_pool_forward_train(ctx, 0.9, X, S, B, rm1, rv1, **kw)
    # (a) the separate entry points
    rm2, rv2 = dev(np.zeros(c, np.float32)), dev(np.ones(c, np.float32))
    if prestats:
        y2, sm2, si2 = zb.batch_norm_2d_forward_train_prestats(ctx, 0.9, X, S, B, rm2, rv2, partial, rows, shift, relu=True)
    else:
        y2, sm2, si2 = zb.batch_norm_2d_forward_train(ctx, 0.9, X, S, B, rm2, rv2, layout=ZB_NHWC, relu=True)
    yp2, idx2 = zb.max_pool_2d_indexed(ctx, y2, 3, 2, 1)
    for a, b in ((yp, yp2), (idx, idx2), (sm, sm2), (si, si2), (rm1, rm2), (rv1, rv2)):
        np.testing.assert_array_equal(host(a), host(b))
    dyp = rng.standard_normal(tuple(yp.shape)).astype(np.float32)
    DYP = dev(dyp)
    dx, ds, db = zb.batch_norm_relu_max_pool_backward(ctx, X, DYP, idx, S, B, sm, si)
    dy2 = zb.max_pool_2d_indexed_backward(ctx, DYP, idx2, tuple(y2.shape), 3, 2, 1)
    dx2, ds2, db2 = zb.batch_norm_2d_relu_backward(ctx, X, dy2, S, B, sm2, si2, layout=ZB_NHWC)
    assert rel_err(host(dx), host(dx2)) < 2e-5 and rel_err(host(ds), host(ds2)) < 2e-5 and rel_err(host(db), host(db2)) < 2e-5
    # (b) the oracle's separate nodes
    bn_ref, _, _, sm_ref, si_ref = zo.bn2d_fwd_train(x, scale, bias, np.zeros(c, np.float32), np.ones(c, np.float32), 0.9)
    act = zo.relu(bn_ref)
    yp_ref = zo.maxpool2d_fwd(act, 3, 2, 1)
    assert rel_err(nchw(host(yp)), yp_ref) < 2e-5
    g_act = zo.maxpool2d_bwd(act, nchw(dyp), 3, 2, 1)
    g_bn = zo.ewise("mul", g_act, (bn_ref > 0).astype(np.float32))
    dx_ref, ds_ref, db_ref = zo.bn2d_bwd(x, g_bn, scale, sm_ref, si_ref)
    assert rel_err(nchw(host(dx)), dx_ref) < 2e-4 and rel_err(host(ds), ds_ref) < 2e-4 and rel_err(host(db), db_ref) < 2e-4
    ctx.check()


def test_bn_relu_maxpool_unsupported_geometries(zb, ctx):
    """Anything but f32 NHWC 3x3 / 2 / 1 with C / 4 a power of two reports ZB_ERR_UNSUPPORTED (the caller composes the separate calls)."""
    from zenu_b200 import ZenuB200Error
    x = torch.randn((2, 6, 6, 24), device="cuda")   # C / 4 = 6: not a power of two
    c = 24
    args = (dev(np.ones(c, np.float32)), dev(np.zeros(c, np.float32)), dev(np.zeros(c, np.float32)), dev(np.ones(c, np.float32)))
    with pytest.raises(ZenuB200Error):
        zb.batch_norm_relu_max_pool_forward_train(ctx, 0.9, x, *args)
    x = torch.randn((2, 6, 6, 16), device="cuda")
    args = tuple(dev(a) for a in (np.ones(16, np.float32), np.zeros(16, np.float32), np.zeros(16, np.float32), np.ones(16, np.float32)))
    with pytest.raises(ZenuB200Error):
        zb.batch_norm_relu_max_pool_forward_train(ctx, 0.9, x, *args, kernel=2, stride=2, pad=0)


# ------------------------------------------------------------------------------------------------ elementwise / pool / loss / optim
def test_elementwise_vs_oracle(zb, ctx):
    rng = np.random.default_rng(5)
    for n in (1, 7, 1024, 100003):
        a = rng.standard_normal(n).astype(np.float32)
        b = (rng.standard_normal(n) + 3.0).astype(np.float32)
        for op in ("add", "sub", "mul", "div"):
            np.testing.assert_array_equal(host(zb.binary(ctx, op, dev(a), dev(b))), zo.ewise(op, a, b))
            np.testing.assert_allclose(host(zb.binary(ctx, op, dev(a), 1.5)), zo.ewise(op, a, 1.5), rtol=1e-6)
        np.testing.assert_array_equal(host(zb.relu(ctx, dev(a))), zo.relu(a))
        np.testing.assert_array_equal(host(zb.relu_backward_mask(ctx, dev(a))), zo.relu_backward_mask(a))
        np.testing.assert_array_equal(host(zb.relu_bwd(ctx, dev(a), dev(b))), zo.ewise("mul", b, zo.relu_backward_mask(a)))
    a = rng.standard_normal((37, 24)).astype(np.float32)
    v = rng.standard_normal(24).astype(np.float32)
    np.testing.assert_array_equal(host(zb.binary(ctx, "add", dev(a), dev(v))), a + v)
    np.testing.assert_allclose(host(zb.sum_rows(ctx, dev(a))), a.astype(np.float64).sum(0), rtol=1e-5, atol=1e-5)
    x = rng.standard_normal((3, 5, 4, 6)).astype(np.float32)
    np.testing.assert_array_equal(host(zb.to_nhwc(ctx, dev(x))), nhwc(x))
    np.testing.assert_array_equal(host(zb.to_nchw(ctx, dev(nhwc(x)))), x)
    e = torch.empty(0, device="cuda")
    assert zb.relu(ctx, e).numel() == 0  # empty input


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_pool_and_gap(zb, ctx, layout):
    from zenu_b200 import ZB_NCHW, ZB_NHWC
    rng = np.random.default_rng(8)
    x = np.maximum(rng.standard_normal((2, 6, 11, 9)), 0).astype(np.float32)   # post-ReLU input with ties at 0
    y_ref = zo.maxpool2d_fwd(x, 3, 2, 1)
    dy = rng.standard_normal(y_ref.shape).astype(np.float32)
    dx_ref = zo.maxpool2d_bwd(x, dy, 3, 2, 1)
    if layout == "nchw":
        L, cv, back = ZB_NCHW, dev, host
    else:
        L, cv, back = ZB_NHWC, (lambda a: dev(nhwc(a))), (lambda t: nchw(host(t)))
    y = zb.max_pool_2d(ctx, cv(x), 3, 2, 1, layout=L)
    np.testing.assert_array_equal(back(y), y_ref)
    dx = zb.max_pool_2d_backward(ctx, cv(x), cv(dy), 3, 2, 1, layout=L)
    np.testing.assert_allclose(back(dx), dx_ref, rtol=1e-6, atol=1e-6)
    g = zb.global_avg_pool(ctx, cv(x), layout=L)
    np.testing.assert_allclose(host(g), zo.gap_fwd(x), rtol=1e-6, atol=1e-7)
    dg = rng.standard_normal((2, 6)).astype(np.float32)
    np.testing.assert_allclose(back(zb.global_avg_pool_backward(ctx, dev(dg), cv(x).shape, layout=L)), zo.gap_bwd(dg, (11, 9)), rtol=1e-6)


def test_input_pipeline(zb, ctx):
    """uint8 batch -> normalised NCHW float batch and int labels -> one-hot, against numpy."""
    from zenu_b200 import ZB_NCHW, ZB_NHWC
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (3, 10, 7, 3), dtype=np.uint8)          # NHWC as decoded
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    ref = ((img.astype(np.float64) / 255.0 - np.array(mean)) / np.array(std)).transpose(0, 3, 1, 2)
    for dt, tol in ((torch.float32, 1e-6), (torch.float64, 1e-6)):
        got = zb.input_u8_to_float(ctx, torch.from_numpy(img).cuda(), mean, std, src_layout=ZB_NHWC, dtype=dt)
        assert got.shape == (3, 3, 10, 7) and maxabs(host(got), ref) < tol * 10
        got2 = zb.input_u8_to_float(ctx, torch.from_numpy(np.ascontiguousarray(img.transpose(0, 3, 1, 2))).cuda(), mean, std, src_layout=ZB_NCHW, dtype=dt)
        np.testing.assert_array_equal(host(got2), host(got))
    plain = zb.input_u8_to_float(ctx, torch.from_numpy(img).cuda())
    assert maxabs(host(plain), img.transpose(0, 3, 1, 2) / 255.0) < 1e-6
    labels = rng.integers(0, 10, 17).astype(np.int32)
    oh = host(zb.onehot(ctx, torch.from_numpy(labels).cuda(), 10))
    np.testing.assert_array_equal(oh, np.eye(10, dtype=np.float32)[labels])
    ctx.check()


def test_softmax_xent(zb, ctx):
    rng = np.random.default_rng(21)
    z = (rng.standard_normal((16, 1000)) * 3).astype(np.float32)
    t = np.zeros((16, 1000), np.float32)
    t[np.arange(16), rng.integers(0, 1000, 16)] = 1.0
    loss_ref, dz_ref = zo.softmax_xent(z.astype(np.float64), t.astype(np.float64))
    loss, dz = zb.softmax_cross_entropy(ctx, dev(z), dev(t))
    assert abs(float(host(loss)[0]) - loss_ref) < 1e-5 * max(1.0, abs(loss_ref))
    assert rel_err(host(dz), dz_ref) < 1e-5


def test_optimizers_vs_oracle(zb, ctx, lit):
    rng = np.random.default_rng(31)
    n = 10007
    p0 = rng.standard_normal(n).astype(np.float32)
    g = rng.standard_normal(n).astype(np.float32)
    p_ref = p0.copy()
    zo.sgd_step(p_ref, g, 0.01)
    p = dev(p0)
    zb.sgd_step(ctx, p, dev(g), 0.01)
    np.testing.assert_allclose(host(p), p_ref, rtol=1e-6, atol=1e-7)
    for decay in (False, True):
        p_ref, m_ref, v_ref = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
        p, m, v = dev(p0), dev(np.zeros(n, np.float32)), dev(np.zeros(n, np.float32))
        for step in (1, 2, 3):
            zo.adam_step(p_ref, g, m_ref, v_ref, 0.01, 0.9, 0.999, 1e-8, step, 0.01, decay)
            zb.adam_step(ctx, p, dev(g), m, v, 0.01, 0.9, 0.999, 1e-8, step, 0.01, decay)
        np.testing.assert_allclose(host(p), p_ref, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(host(m), m_ref, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------------ compat shims
def test_kernel_sys_shim(zb, ctx):
    """The zenu-cuda-kernel-sys symbol names, called with raw device pointers (strided and unit-stride)."""
    from zenu_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(2)
    a = rng.standard_normal(64).astype(np.float32)
    b = rng.standard_normal(64).astype(np.float32)
    A, B = dev(a), dev(b)
    C = torch.zeros(64, device="cuda")
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.array_array_add_float.argtypes = [vp, ci, vp, ci, vp, ci, ci]
    lib.array_array_add_float.restype = None
    lib.array_array_add_float(A.data_ptr(), 1, B.data_ptr(), 1, C.data_ptr(), 1, 64)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(host(C), a + b)
    C.zero_()
    lib.array_array_add_float(A.data_ptr(), 2, B.data_ptr(), 2, C.data_ptr(), 1, 32)   # strided inputs
    torch.cuda.synchronize()
    np.testing.assert_array_equal(host(C)[:32], a[::2] + b[::2])
    lib.relu_float.argtypes = [vp, vp, cf, ci, ci, ci]
    lib.relu_float.restype = None
    lib.relu_float(A.data_ptr(), C.data_ptr(), 0.0, 64, 1, 1)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(host(C), np.maximum(a, 0))
    dy = rng.standard_normal((3, 4, 5, 6)).astype(np.float32)
    DB = torch.zeros(4, device="cuda")
    lib.conv2d_bias_bkwd_float.argtypes = [vp, vp, ci, ci, ci, ci]
    lib.conv2d_bias_bkwd_float.restype = None
    lib.conv2d_bias_bkwd_float(dev(dy).data_ptr(), DB.data_ptr(), 3, 4, 5, 6)
    torch.cuda.synchronize()
    np.testing.assert_allclose(host(DB), zo.conv2d_bias_bkwd(dy), rtol=1e-5)
    out = ctypes.c_float(0)
    lib.memory_access_float.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_float)]
    lib.memory_access_float.restype = None
    lib.memory_access_float(A.data_ptr(), 5, ctypes.byref(out))
    assert out.value == a[5]


def test_cudnn_frontend_shim(zb, ctx, golden_dir):
    """create/check/workspace/execute on the reference's conv fixture through the cuDNN-frontend-shaped ABI."""
    from zenu_b200 import _lib
    lib = _lib.load()

    class Shape(ctypes.Structure):
        _fields_ = [("num_dims", ctypes.c_size_t), ("dims", ctypes.c_int64 * 8), ("strides", ctypes.c_int64 * 8)]

    class Info(ctypes.Structure):
        _fields_ = [("padding", ctypes.c_int64 * 2), ("stride", ctypes.c_int64 * 2), ("dilation", ctypes.c_int64 * 2),
                    ("num_dims", ctypes.c_int64)]

    class Bufs(ctypes.Structure):
        _fields_ = [("X", ctypes.c_void_p), ("filter", ctypes.c_void_p), ("Y", ctypes.c_void_p)]

    def shp(dims):
        s = Shape()
        s.num_dims = 4
        st = 1
        for i in (3, 2, 1, 0):
            s.dims[i] = dims[i]
            s.strides[i] = st
            st *= dims[i]
        return s

    d = np.load(os.path.join(golden_dir, "conv2d.npz"))
    x, w = dev(d["input"]), dev(d["filter"])
    y = torch.zeros(d["output"].shape, device="cuda")
    info = Info()
    info.padding[:] = [1, 1]; info.stride[:] = [1, 1]; info.dilation[:] = [1, 1]; info.num_dims = 2
    desc = ctypes.c_void_p()
    xs, ws, ys = shp(x.shape), shp(w.shape), shp(y.shape)
    lib.create_conv_descriptor.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 4
    assert lib.create_conv_descriptor(ctypes.byref(desc), 1, ctypes.byref(xs), ctypes.byref(ws), ctypes.byref(ys), ctypes.byref(info)) == 0
    lib.check_conv_graph.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    assert lib.check_conv_graph(desc, None) == 0
    size = ctypes.c_int64(-1)
    lib.get_conv_workspace_size.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    assert lib.get_conv_workspace_size(desc, ctypes.byref(size)) == 0 and size.value == 0
    bufs = Bufs(x.data_ptr(), w.data_ptr(), y.data_ptr())
    lib.execute_conv_forward.argtypes = [ctypes.c_void_p] * 4
    assert lib.execute_conv_forward(desc, ctypes.byref(bufs), None, None) == 0
    torch.cuda.synchronize()
    assert maxabs(host(y), d["output"]) < 5e-3   # the shim runs the ctx default math (TF32 tensor cores, like cuDNN's default for FLOAT convs)
    lib.destroy_conv_descriptor.argtypes = [ctypes.c_void_p]
    lib.destroy_conv_descriptor(desc)


def test_input_stage(zb, ctx):
    """zb_input_stage_* (SURVEY 8f-3): pinned multi-buffered uint8 staging + copy stream + device expansion.  Every batch that goes in
    comes out as (u8 / 255 - mean) / std in NCHW and the labels as one-hot rows, slots reused several times with all the copies in
    flight ahead of the waits."""
    from zenu_b200 import ZB_NCHW, ZB_NHWC, nn
    rng = np.random.default_rng(9)
    n, c, h, w, classes = 5, 3, 12, 9, 7
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    for layout in (ZB_NHWC, ZB_NCHW):
        st = nn.InputStage(ctx, n, c, h, w, classes, mean=mean, std=std, slots=3, src_layout=layout)
        assert st.h2d_bytes == n * c * h * w + 4 * n
        sent = {}
        outs = []
        for i in range(8):                      # keep two submits ahead of the consumer
            if i < 8:
                img, lab = st.host_buffers(i % 3)
                img[...] = rng.integers(0, 256, img.shape, dtype=np.uint8)
                lab[...] = rng.integers(0, classes, n, dtype=np.int32)
                sent[i] = (img.copy(), lab.copy())
                st.submit(i % 3)
            if i >= 2:
                x, t = st.wait((i - 2) % 3)
                outs.append((i - 2, x.clone(), t.clone()))
        for j in (6, 7):
            x, t = st.wait(j % 3)
            outs.append((j, x.clone(), t.clone()))
        ctx.check()
        for j, x, t in outs:
            img, lab = sent[j]
            chw = img.transpose(0, 3, 1, 2) if layout == ZB_NHWC else img
            ref = (chw.astype(np.float64) / 255.0 - np.array(mean)[None, :, None, None]) / np.array(std)[None, :, None, None]
            assert x.shape == (n, c, h, w) and maxabs(host(x), ref) < 1e-5, j
            np.testing.assert_array_equal(host(t), np.eye(classes, dtype=np.float32)[lab])
        with pytest.raises(Exception):
            st.wait(0)                          # no submit pending on that slot
        st.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_reductions_and_strided_copy_vs_oracle(zb, ctx, dtype):
    """SURVEY 8a-12 / 8a-13 as first-class ops through the C ABI: Matrix::sum(axis) / sum_to / mean / variance
    (zenu-matrix/src/operation/{sum.rs:9-92, mean.rs:8-20, var.rs:18-26}), zeros / copy / strided copy_from
    (device/mod.rs:33-69, operation/copy_from.rs:9-123), against the oracle's restatements (reference accumulation order)."""
    rng = np.random.default_rng(17)
    tol = 2e-5 if dtype == np.float32 else 1e-12
    lit = np.arange(120, dtype=dtype).reshape(2, 3, 4, 5)           # the reference's own test tensor (sum.rs:104-160)
    for axis in range(4):
        for keep in (False, True):
            got = host(zb.sum_axis(ctx, dev(lit), axis, keep))
            np.testing.assert_array_equal(got, zo.sum_axis(lit, axis, keep))
    shapes = [((7, 5, 3), None), ((64, 1000), None), ((3, 700, 33), None), ((1, 5000, 1), None), ((6, 2, 4, 130), None), ((2048, 9), None),
              ((4, 100000), None)]
    for shape, _ in shapes:
        a = (rng.standard_normal(shape) + 0.5).astype(dtype)
        A = dev(a)
        for axis in range(len(shape)):
            ref = zo.sum_axis(a.astype(np.float64), axis)
            scale = np.abs(a).astype(np.float64).sum(axis).max() + 1e-30
            assert np.max(np.abs(host(zb.sum_axis(ctx, A, axis)) - ref)) / scale < tol, (shape, axis)
            np.testing.assert_allclose(host(zb.mean_axis(ctx, A, axis, True)), zo.mean_axis(a.astype(np.float64), axis, True), rtol=50 * tol,
                                       atol=50 * tol)
            var, mean = zb.variance_axis(ctx, A, axis, want_mean=True)
            np.testing.assert_allclose(host(var), zo.variance_axis(a.astype(np.float64), axis), rtol=100 * tol, atol=10 * tol)
            np.testing.assert_allclose(host(mean), zo.mean_axis(a.astype(np.float64), axis), rtol=50 * tol, atol=50 * tol)
            np.testing.assert_array_equal(host(zb.variance_axis(ctx, A, axis)), host(var))      # mean_out = NULL path
    assert abs(float(host(zb.variance_axis(ctx, dev(np.array([1, 2, 3, 4], dtype)), 0))) - 1.25) < 1e-6   # var.rs:35-40
    # sum_to: bias / gamma / beta gradients and broadcast backward (functions/sum_to.rs)
    g = rng.standard_normal((5, 6, 7, 8)).astype(dtype)
    G = dev(g)
    for target in ((8,), (7, 8), (1, 8), (6, 1, 8), (5, 1, 7, 1), (1, 1, 1, 1), (5, 6, 7, 8), (1, 6, 1, 1), ()):
        got = host(zb.sum_to(ctx, G, target))
        ref = zo.sum_to(g.astype(np.float64), target)
        assert got.shape == tuple(target)
        np.testing.assert_allclose(got, ref, rtol=50 * tol, atol=200 * tol)
    from zenu_b200 import ZenuB200Error
    with pytest.raises(ZenuB200Error):
        zb.sum_to(ctx, G, (5, 6, 7, 3))
    e = torch.empty((4, 0, 3), dtype=G.dtype, device="cuda")
    np.testing.assert_array_equal(host(zb.sum_axis(ctx, e, 1)), np.zeros((4, 3), dtype))      # empty axis -> zeros
    # fill / copy / strided copies
    x = torch.empty(100003, dtype=G.dtype, device="cuda")
    assert float(zb.fill(ctx, x, 0.0).abs().max()) == 0.0 and float(zb.fill(ctx, x, 2.5).min()) == 2.5
    np.testing.assert_array_equal(host(zb.copy(ctx, G)), g)
    t4 = dev(rng.standard_normal((3, 4, 5, 6)).astype(dtype))
    cases = [t4.permute(0, 2, 3, 1), t4.permute(3, 2, 1, 0), t4[:, ::2, 1:4, ::3], t4[1], t4.transpose(1, 2)[..., 2],
             t4[:1, :1, :1, :].expand(3, 4, 5, 6)]                                      # transposes, stepped slices, a broadcast source
    for v in cases:
        out = torch.full(tuple(v.shape), -7.0, dtype=v.dtype, device="cuda")
        zb.copy_strided(ctx, v, out)
        np.testing.assert_array_equal(host(out), zo.copy_strided(host(v)))
    big = torch.zeros((6, 10), dtype=G.dtype, device="cuda")
    zb.copy_strided(ctx, t4[0, 0], big[1:, 2:8])                                        # strided DESTINATION (assign into a slice)
    ref = np.zeros((6, 10), dtype)
    ref[1:, 2:8] = host(t4[0, 0])
    np.testing.assert_array_equal(host(big), ref)
    ctx.check()


@pytest.mark.parametrize("case", [(2, 64, 8, 8, 128), (3, 96, 6, 10, 40), (2, 256, 14, 14, 64), (5, 32, 4, 8, 300), (2, 128, 28, 28, 512)])
@pytest.mark.parametrize("math", ["tf32", "tf32x3"])
def test_nchw_pointwise_conv_without_staging(zb, ctx, case, math):
    """The reference contract (NCHW / KCRS, zenu-matrix/src/nn/conv/interface.rs:270-281) for 1x1 / stride-1 convs is served by
    batched per-image GEMMs straight on the NCHW tensors (UmmaParams::batch_mode): no NHWC staging copies.  Ragged output-channel
    tiles (rows of the NEXT image land in masked accumulator rows), ragged pixel tiles, bias, all three passes, vs the oracle; the
    plan description must show the batched launch and no transpose."""
    from zenu_b200 import ZB_NCHW
    n, c, h, w, k = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    wt = (rng.standard_normal((k, c, 1, 1)) * np.sqrt(2.0 / c)).astype(np.float32)
    b = rng.standard_normal(k).astype(np.float32)
    y_ref = zo.conv2d_bias_add(zo.conv2d_fwd(x.astype(np.float64), wt.astype(np.float64), 0, 1, 1), b.astype(np.float64))
    dy = rng.standard_normal(y_ref.shape).astype(np.float32)
    m = math_of(zb, math)
    for op in (zb.PLAN_FPROP, zb.PLAN_DGRAD, zb.PLAN_WGRAD):
        plan = zb.conv_plan_describe(ctx, op, x.shape, wt.shape, layout=ZB_NCHW, math=m)
        if op == zb.PLAN_DGRAD and k % 32 != 0:     # K blocks of the dgrad walk the output channels: 40 of them is not a whole number
            assert "batch_mode" not in plan          # of blocks, the next image's rows would enter the sum -> the FFMA kernel serves it
            continue
        assert "batch_mode=" + ("2" if op == zb.PLAN_WGRAD else "1") in plan and "transpose" not in plan, plan
    X, W, DY = dev(x), dev(wt), dev(dy)
    tol = TOL[math]
    assert rel_err(host(zb.conv_fwd(ctx, X, W, 0, 1, 1, bias=dev(b), layout=ZB_NCHW, math=m)), y_ref) < tol
    dx = zb.conv_bkwd_data(ctx, DY, W, X.shape, 0, 1, 1, layout=ZB_NCHW, math=m)
    assert rel_err(host(dx), zo.conv2d_bkwd_data(dy.astype(np.float64), wt.astype(np.float64), x.shape, 0, 1, 1)) < tol
    dw = zb.conv_bkwd_weight(ctx, DY, X, W.shape, 0, 1, 1, layout=ZB_NCHW, math=m)
    assert rel_err(host(dw), zo.conv2d_bkwd_filter(dy.astype(np.float64), x.astype(np.float64), wt.shape, 0, 1, 1)) < tol
    ctx.check()


@pytest.mark.parametrize("dtype,math", [(np.float32, "fp32"), (np.float64, "fp32"), (np.float32, "tf32"), (np.float32, "tf32x3")])
def test_wgrad_is_run_to_run_deterministic(zb, ctx, dtype, math):
    """Every wgrad path reduces its split-K partials in a fixed order (the FFMA / DFMA kernels used atomicAdd in round 1): the same
    call twice gives the same bits, for the SIMT kernels (ZB_MATH_FP32, all of f64) and the tensor-core kernels alike."""
    from zenu_b200 import ZB_NHWC
    rng = np.random.default_rng(41)
    for (n, c, h, k, r, pad, stride) in ((16, 32, 20, 48, 3, 1, 1), (8, 3, 33, 16, 5, 2, 2), (4, 64, 14, 64, 1, 0, 1)):
        x = rng.standard_normal((n, h, h, c)).astype(dtype)
        p = (h + 2 * pad - r) // stride + 1
        dy = rng.standard_normal((n, p, p, k)).astype(dtype)
        X, DY = dev(x), dev(dy)
        outs = [host(zb.conv_bkwd_weight(ctx, DY, X, (k, r, r, c), pad, stride, 1, layout=ZB_NHWC, math=math_of(zb, math))) for _ in range(3)]
        np.testing.assert_array_equal(outs[0], outs[1])
        np.testing.assert_array_equal(outs[0], outs[2])
        ref = zo.conv2d_bkwd_filter(nchw(dy).astype(np.float64), nchw(x).astype(np.float64), (k, c, r, r), pad, stride, 1)
        assert rel_err(nchw(outs[0]), ref) < (1e-10 if dtype == np.float64 else TOL[math])
    ctx.check()
