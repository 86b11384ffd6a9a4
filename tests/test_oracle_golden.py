"""Pins the CPU oracle (oracle/) against every golden vector the reference's own tests hold for the
hot path.  Tolerances are the reference tests' own (max-abs-diff,
zenu-test/src/lib.rs:3-24)."""
import json
import os

import numpy as np
import pytest

from oracle import zenu_oracle as zo


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64).ravel() - np.asarray(b, np.float64).ravel())))


@pytest.fixture(scope="module")
def lit(golden_dir):
    with open(os.path.join(golden_dir, "literals.json")) as f:
        return json.load(f)


@pytest.fixture(params=["plain", "openblas"])
def gemm_backend(request):
    if request.param == "openblas":
        if not zo.use_openblas():
            pytest.skip("OpenBLAS not found")
    else:
        zo.use_plain_gemm()
    yield request.param
    zo.use_plain_gemm()


def test_conv2d_json(golden_dir, gemm_backend):
    # zenu-matrix/src/nn/conv/mod.rs:131-179 (tol 1e-4); fixture has dy = ones
    d = np.load(os.path.join(golden_dir, "conv2d.npz"))
    x, w = d["input"], d["filter"]
    y = zo.conv2d_fwd(x, w, pad=1, stride=1, dil=1)
    assert maxabs(y, d["output"]) < 1e-4
    dy = np.ones_like(d["output"])
    assert maxabs(zo.conv2d_bkwd_data(dy, w, x.shape, 1, 1, 1), d["grad_input"]) < 1e-4
    assert maxabs(zo.conv2d_bkwd_filter(dy, x, w.shape, 1, 1, 1), d["grad_weight"]) < 1e-4


def test_conv_bias_json(golden_dir, gemm_backend):
    # zenu-autograd/src/nn/conv/conv_with_bias.rs:115-152
    d = np.load(os.path.join(golden_dir, "conv_bias.npz"))
    x, w, b = d["input"], d["filter"], d["bias"]
    y = zo.conv2d_bias_add(zo.conv2d_fwd(x, w, 1, 1, 1), b)
    assert maxabs(y, d["output"]) < 1e-4
    dy = np.ones_like(y)
    assert maxabs(zo.conv2d_bkwd_data(dy, w, x.shape, 1, 1, 1), d["grad_input"]) < 1e-4
    assert maxabs(zo.conv2d_bkwd_filter(dy, x, w.shape, 1, 1, 1), d["grad_weight"]) < 1e-4
    assert maxabs(zo.conv2d_bias_bkwd(dy), d["grad_bias"]) < 1e-4


def test_conv_fwd_small_literal(lit, gemm_backend):
    # zenu-cuda/src/cudnn/graph_conv.rs:433-623
    c = lit["conv_fwd_small"]
    x = np.array(c["input"], np.float32).reshape(c["x_shape"])
    w = np.array(c["filter"], np.float32).reshape(c["w_shape"])
    y = zo.conv2d_fwd(x, w, c["pad"], c["stride"], c["dilation"])
    assert maxabs(y, c["output"]) < 1e-5


def test_bn_forward_train(lit):
    c = lit["bn_fwd_train"]
    x = np.array(c["x"], np.float32).reshape(c["shape"])
    z = np.zeros(2, np.float32)
    y, rm, rv, sm, si = zo.bn2d_fwd_train(x, np.array(c["scale"], np.float32), np.array(c["bias"], np.float32), z, z,
                                           c["momentum"])
    tol = c["tol"]
    assert maxabs(y, c["y"]) < tol
    assert maxabs(rm, c["running_mean"]) < tol
    assert maxabs(rv, c["running_variance"]) < tol
    assert maxabs(sm, c["saved_mean"]) < tol
    assert maxabs(si, c["saved_inv_std"]) < tol


def test_bn_backward(lit):
    c = lit["bn_bwd"]
    x = np.array(c["x"], np.float32).reshape(c["shape"])
    dy = np.array(c["y_grad"], np.float32).reshape(c["shape"])
    dx, ds, db = zo.bn2d_bwd(x, dy, np.array(c["scale"], np.float32), np.array(c["saved_mean"], np.float32),
                             np.array(c["saved_inv_std"], np.float32))
    assert maxabs(dx, c["x_grad"]) < c["tol"]
    assert maxabs(ds, c["scale_grad"]) < c["tol"]
    assert maxabs(db, c["bias_grad"]) < c["tol"]


def test_bn_inference(lit):
    c = lit["bn_infer"]
    x = np.array(c["x"], np.float32).reshape(c["shape"])
    y = zo.bn2d_fwd_infer(x, np.array(c["scale"], np.float32), np.array(c["bias"], np.float32),
                          np.array(c["mean"], np.float32), np.array(c["variance"], np.float32))
    assert maxabs(y, c["y"]) < c["tol"]


def test_bn_autograd_case(lit):
    # zenu-autograd/src/nn/batch_norm.rs:284-737: y = BN(x); (y * y_grad).backward()
    c = lit["bn_autograd"]
    x = np.array(c["x"], np.float32).reshape(c["shape"])
    dy = np.array(c["y_grad"], np.float32).reshape(c["shape"])
    scale, bias = np.array(c["scale"], np.float32), np.array(c["bias"], np.float32)
    y, rm, rv, sm, si = zo.bn2d_fwd_train(x, scale, bias, np.array(c["prev_mean"], np.float32),
                                           np.array(c["prev_var"], np.float32), c["momentum"])
    assert maxabs(y, c["y"]) < c["tol_y"]
    dx, ds, db = zo.bn2d_bwd(x, dy, scale, sm, si)
    assert maxabs(dx, c["x_grad"]) < c["tol_x_grad"]
    assert maxabs(ds, c["scale_grad"]) < c["tol_param_grad"]
    assert maxabs(db, c["bias_grad"]) < c["tol_param_grad"]
    # saved stats == None path recomputes from x (batch_norm.rs:355-368)
    dx2, _, _ = zo.bn2d_bwd(x, dy, scale)
    assert maxabs(dx2, c["x_grad"]) < c["tol_x_grad"]


def test_gemm_literal(lit, gemm_backend):
    c = lit["gemm_3x4_4x5"]
    a = np.array(c["a"], np.float32).reshape(3, 4)
    b = np.array(c["b"], np.float32).reshape(4, 5)
    out = zo.gemm(a, b)
    assert float(np.abs(out.ravel() - np.array(c["c"])).sum()) < c["tol_asum"]
    # transposed operand forms (zenu-cuda/src/cublas/mod.rs:376-566 exercise N/T)
    assert float(np.abs(zo.gemm(np.ascontiguousarray(a.T), b, trans_a=True) - out).sum()) < 1e-5
    assert float(np.abs(zo.gemm(a, np.ascontiguousarray(b.T), trans_b=True) - out).sum()) < 1e-5


def test_relu_literal(lit):
    c = lit["relu"]
    x = np.array(c["x"], np.float32)
    assert maxabs(zo.relu(x), c["y"]) < c["tol"]
    assert maxabs(zo.relu_backward_mask(x), c["mask"]) < c["tol"]


def _mlp_step(net, optimizer, state):
    """zenu-optimizer/tests/net_test.rs:94-106: Linear(2,4) -> Linear(4,4), MSE, one update."""
    x = np.array(net["input"], np.float32).reshape(1, 2)
    t = np.array(net["target"], np.float32).reshape(1, 4)
    h = zo.linear_fwd(x, state["w1"], state["b1"])
    o = zo.linear_fwd(h, state["w2"], state["b2"])
    # mse = sum((t - o)^2) / batch, batch = y_true.shape[0] = 4: the target is 1-D [4] in the test (mse.rs:10)
    do = (-2.0 * (t - o) / np.float32(4)).astype(np.float32)
    dh, dw2, db2 = zo.linear_bwd(h, state["w2"], do)
    _, dw1, db1 = zo.linear_bwd(x, state["w1"], dh)
    optimizer(state, {"w1": dw1, "b1": db1, "w2": dw2, "b2": db2})


def _fresh(net):
    return {"w1": np.array(net["w1"], np.float32).reshape(4, 2), "b1": np.array(net["b1"], np.float32),
            "w2": np.array(net["w2"], np.float32).reshape(4, 4), "b2": np.array(net["b2"], np.float32)}


def _check(state, c):
    for k in ("w1", "b1", "w2", "b2"):
        assert maxabs(state[k], c[k]) < c["tol"], k


def test_sgd_literal(lit):
    c = lit["optim_net"]
    st = _fresh(c["net"])
    _mlp_step(c["net"], lambda s, g: [zo.sgd_step(s[k], g[k], c["sgd"]["lr"]) for k in s], st)
    _check(st, c["sgd"])


def test_adam_literal(lit):
    c = lit["optim_net"]
    a = c["adam"]
    st = _fresh(c["net"])
    for _ in range(2):  # the test builds a fresh Adam (m=v=0, step=0) before each of the two steps
        m = {k: np.zeros_like(v) for k, v in st.items()}
        v = {k: np.zeros_like(vv) for k, vv in st.items()}
        _mlp_step(c["net"], lambda s, g: [zo.adam_step(s[k], g[k], m[k], v[k], a["lr"], a["beta1"], a["beta2"],
                                                        a["eps"], 1) for k in s], st)
    _check(st, a)


def test_adamw_literal(lit):
    c = lit["optim_net"]
    a = c["adamw"]
    st = _fresh(c["net"])
    m = {k: np.zeros_like(v) for k, v in st.items()}
    v = {k: np.zeros_like(vv) for k, vv in st.items()}
    for step in (1, 2):
        _mlp_step(c["net"], lambda s, g: [zo.adam_step(s[k], g[k], m[k], v[k], a["lr"], a["beta1"], a["beta2"],
                                                        a["eps"], step, a["weight_decay"], k.startswith("w"))
                                          for k in s], st)
    _check(st, a)


def test_reductions_reference_literals():
    """The reference's own literals for Matrix::sum / variance (zenu-matrix/src/operation/sum.rs:104-160 test_4d: arange(120) as
    [2,3,4,5]; operation/var.rs:35-40: variance([1,2,3,4]) = 1.25) and sum_to's right-aligned semantics."""
    src = np.arange(120, dtype=np.float32).reshape(2, 3, 4, 5)
    s0 = zo.sum_axis(src, 0)
    assert s0.shape == (3, 4, 5)
    np.testing.assert_array_equal(s0.ravel(), np.arange(60, 179, 2, dtype=np.float32))           # sum.rs:124-133
    s1 = zo.sum_axis(src, 1)
    assert s1.shape == (2, 4, 5)
    np.testing.assert_array_equal(s1.ravel()[:20], np.arange(60, 118, 3, dtype=np.float32))      # sum.rs:135-141
    assert zo.sum_axis(src, 2).shape == (2, 3, 5) and zo.sum_axis(src, 3).shape == (2, 3, 4)
    for axis in range(4):
        np.testing.assert_array_equal(zo.sum_axis(src, axis), src.sum(axis))
        np.testing.assert_array_equal(zo.sum_axis(src, axis, keep_dim=True), src.sum(axis, keepdims=True))
    assert abs(float(zo.variance_axis(np.array([1.0, 2.0, 3.0, 4.0], np.float32), 0)) - 1.25) < 1e-6
    np.testing.assert_array_equal(zo.sum_to(src, (4, 5)), src.sum((0, 1)))
    np.testing.assert_array_equal(zo.sum_to(src, (2, 1, 4, 1)), src.sum((1, 3), keepdims=True))
    np.testing.assert_array_equal(zo.sum_to(src, (2, 3, 4, 5)), src)
    np.testing.assert_allclose(zo.mean_axis(src, 1), src.mean(1), rtol=1e-6)
