"""ctypes/numpy face of the CPU oracle (oracle/zenu_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (zenu_b200/) never imports this.

All tensors are NCHW / row-major numpy arrays exactly as the reference's `Matrix` default stride
(zenu-matrix/src/dim/mod.rs:70-92); dtype float32 or float64 selects the _f32 / _f64 entry points.
"""
import ctypes
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

I64 = ctypes.c_int64


def build(force=False):
    so = os.path.join(_HERE, "libzenu_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("zenu_oracle.c", "zenu_oracle_impl.inc")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libzenu_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libzenu_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.zo_conv_dim_out_size.restype = I64
        _LIB.zo_conv_dim_out_size.argtypes = [I64] * 5
    return _LIB


def find_openblas():
    """numpy's bundled ILP64 OpenBLAS (same library family the reference links, zenu-matrix/Cargo.toml:12-13)."""
    cands = glob.glob(os.path.join(os.path.dirname(np.__file__), "..", "numpy.libs", "libscipy_openblas64_*.so"))
    return os.path.abspath(cands[0]) if cands else None


def use_openblas(threads=None):
    """Route zo_gemm through OpenBLAS (multithreaded).  Returns True when loaded."""
    p = find_openblas()
    if p is None:
        return False
    ok = lib().zo_set_blas(p.encode()) == 0
    if ok and threads:
        lib().zo_set_blas_threads(int(threads))
    return ok


def use_plain_gemm():
    lib().zo_unset_blas()


def _sfx(a):
    if a.dtype == np.float32:
        return "_f32", ctypes.c_float
    if a.dtype == np.float64:
        return "_f64", ctypes.c_double
    raise TypeError("oracle supports float32/float64 only (zenu-matrix/src/num.rs:44-86)")


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _c(a, dtype=None):
    return np.ascontiguousarray(a, dtype=dtype)


def tf32_round(a, mode="rna"):
    """fp32 -> tf32 (10 explicit mantissa bits) -> fp32: what the tensor-core path does to GEMM/conv OPERANDS (products
    and sums stay fp32).  Not part of the reference's algorithm: it models the device's operand rounding so that model-level
    tests can separate "TF32 operand rounding" (expected, bounded) from kernel bugs.
    mode: rna = nearest, ties away from zero (cvt.rna.tf32.f32, the TMA TFLOAT32 load); rz = truncate (what the MMA does
    to raw fp32 bits)."""
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32)
    if mode == "rna":
        r = (u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    elif mode == "rne":
        r = (u + np.uint32(0x0FFF) + ((u >> np.uint32(13)) & np.uint32(1))) & np.uint32(0xFFFFE000)
    elif mode == "rz":
        r = u & np.uint32(0xFFFFE000)
    else:
        raise ValueError(mode)
    out = r.view(np.float32)
    return np.where(np.isfinite(a), out, a)


def conv_out(i, k, pad, stride, dil):
    return int(lib().zo_conv_dim_out_size(i, k, pad, stride, dil))


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _conv_args(n, c, h, w, k, kh, kw, pad, stride, dil):
    (ph, pw), (sh, sw), (dh, dw) = _pair(pad), _pair(stride), _pair(dil)
    return [I64(v) for v in (n, c, h, w, k, kh, kw, ph, pw, sh, sw, dh, dw)]


def conv2d_fwd(x, w, pad=0, stride=1, dil=1):
    x, w = _c(x), _c(w, x.dtype)
    s, _ = _sfx(x)
    n, c, h, wd = x.shape
    k, c2, kh, kw = w.shape
    assert c == c2
    (ph, pw), (sh, sw), (dh, dw) = _pair(pad), _pair(stride), _pair(dil)
    y = np.empty((n, k, conv_out(h, kh, ph, sh, dh), conv_out(wd, kw, pw, sw, dw)), x.dtype)
    rc = getattr(lib(), "zo_conv2d_fwd" + s)(_p(x), _p(w), _p(y), *_conv_args(n, c, h, wd, k, kh, kw, pad, stride, dil))
    assert rc == 0
    return y


def conv2d_bkwd_data(dy, w, x_shape, pad=0, stride=1, dil=1):
    dy, w = _c(dy), _c(w, dy.dtype)
    s, _ = _sfx(dy)
    n, c, h, wd = x_shape
    k, _, kh, kw = w.shape
    dx = np.empty(x_shape, dy.dtype)
    rc = getattr(lib(), "zo_conv2d_bkwd_data" + s)(_p(dy), _p(w), _p(dx), *_conv_args(n, c, h, wd, k, kh, kw, pad, stride, dil))
    assert rc == 0
    return dx


def conv2d_bkwd_filter(dy, x, w_shape, pad=0, stride=1, dil=1):
    dy, x = _c(dy), _c(x, dy.dtype)
    s, _ = _sfx(dy)
    n, c, h, wd = x.shape
    k, _, kh, kw = w_shape
    dw = np.empty(w_shape, dy.dtype)
    rc = getattr(lib(), "zo_conv2d_bkwd_filter" + s)(_p(dy), _p(x), _p(dw), *_conv_args(n, c, h, wd, k, kh, kw, pad, stride, dil))
    assert rc == 0
    return dw


def conv2d_bias_add(x, bias):
    x = _c(x)
    bias = _c(bias, x.dtype).reshape(-1)
    s, _ = _sfx(x)
    n, k, h, w = x.shape
    y = np.empty_like(x)
    getattr(lib(), "zo_conv2d_bias_add" + s)(_p(x), _p(bias), _p(y), I64(n), I64(k), I64(h * w))
    return y


def conv2d_bias_bkwd(dy):
    dy = _c(dy)
    s, _ = _sfx(dy)
    n, k, h, w = dy.shape
    db = np.empty((k,), dy.dtype)
    getattr(lib(), "zo_conv2d_bias_bkwd" + s)(_p(dy), _p(db), I64(n), I64(k), I64(h), I64(w))
    return db


def bn2d_fwd_train(x, scale, bias, run_mean, run_var, momentum):
    """Returns (y, new_run_mean, new_run_var, saved_mean, saved_inv_std)."""
    x = _c(x)
    s, ct = _sfx(x)
    n, c, h, w = x.shape
    scale, bias = _c(scale, x.dtype), _c(bias, x.dtype)
    rm, rv = np.array(run_mean, x.dtype, copy=True), np.array(run_var, x.dtype, copy=True)
    y = np.empty_like(x)
    sm, si = np.empty((c,), x.dtype), np.empty((c,), x.dtype)
    getattr(lib(), "zo_bn2d_fwd_train" + s)(ct(momentum), _p(x), _p(y), _p(scale), _p(bias), _p(rm), _p(rv), _p(sm),
                                            _p(si), I64(n), I64(c), I64(h), I64(w))
    return y, rm, rv, sm, si


def bn2d_bwd(x, dy, scale, saved_mean=None, saved_inv_std=None):
    """Returns (dx, dscale, dbias)."""
    x = _c(x)
    s, _ = _sfx(x)
    dy, scale = _c(dy, x.dtype), _c(scale, x.dtype)
    n, c, h, w = x.shape
    sm = _c(saved_mean, x.dtype) if saved_mean is not None else None
    si = _c(saved_inv_std, x.dtype) if saved_inv_std is not None else None
    dx, ds, db = np.empty_like(x), np.empty((c,), x.dtype), np.empty((c,), x.dtype)
    getattr(lib(), "zo_bn2d_bwd" + s)(_p(x), _p(dy), _p(dx), _p(scale), _p(ds), _p(db), _p(sm), _p(si), I64(n), I64(c),
                                      I64(h), I64(w))
    return dx, ds, db


def bn2d_fwd_infer(x, scale, bias, mean, var):
    x = _c(x)
    s, _ = _sfx(x)
    n, c, h, w = x.shape
    y = np.empty_like(x)
    getattr(lib(), "zo_bn2d_fwd_infer" + s)(_p(x), _p(y), _p(_c(scale, x.dtype)), _p(_c(bias, x.dtype)),
                                            _p(_c(mean, x.dtype)), _p(_c(var, x.dtype)), I64(n), I64(c), I64(h), I64(w))
    return y


def gemm(a, b, trans_a=False, trans_b=False, alpha=1.0, beta=0.0, c=None):
    a = _c(a)
    b = _c(b, a.dtype)
    s, ct = _sfx(a)
    m, k = (a.shape[1], a.shape[0]) if trans_a else a.shape
    k2, n = (b.shape[1], b.shape[0]) if trans_b else b.shape
    assert k == k2
    out = np.zeros((m, n), a.dtype) if c is None else np.array(c, a.dtype, copy=True)
    getattr(lib(), "zo_gemm" + s)(int(trans_a), int(trans_b), I64(m), I64(n), I64(k), ct(alpha), _p(a), I64(a.shape[1]),
                                  _p(b), I64(b.shape[1]), ct(beta), _p(out), I64(n))
    return out


def relu(x, alpha=0.0):
    x = _c(x)
    s, ct = _sfx(x)
    y = np.empty_like(x)
    getattr(lib(), "zo_relu" + s)(_p(x), _p(y), ct(alpha), I64(x.size))
    return y


def relu_backward_mask(x, alpha=0.0):
    x = _c(x)
    s, ct = _sfx(x)
    y = np.empty_like(x)
    getattr(lib(), "zo_relu_backward_mask" + s)(_p(x), _p(y), ct(alpha), I64(x.size))
    return y


_OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3}


def ewise(op, a, b):
    a = _c(a)
    s, ct = _sfx(a)
    out = np.empty_like(a)
    if np.isscalar(b):
        getattr(lib(), "zo_ewise_scalar" + s)(_OPS[op], _p(a), ct(b), _p(out), I64(a.size))
    else:
        b = _c(b, a.dtype)
        assert b.shape == a.shape
        getattr(lib(), "zo_ewise" + s)(_OPS[op], _p(a), _p(b), _p(out), I64(a.size))
    return out


def linear_fwd(x, w, bias=None):
    x = _c(x)
    s, _ = _sfx(x)
    w = _c(w, x.dtype)
    b, i = x.shape
    o = w.shape[0]
    y = np.empty((b, o), x.dtype)
    bias = _c(bias, x.dtype) if bias is not None else None
    getattr(lib(), "zo_linear_fwd" + s)(_p(x), _p(w), _p(bias), _p(y), I64(b), I64(i), I64(o))
    return y


def linear_bwd(x, w, dy):
    """Returns (dx, dw, db)."""
    x = _c(x)
    s, _ = _sfx(x)
    w, dy = _c(w, x.dtype), _c(dy, x.dtype)
    b, i = x.shape
    o = w.shape[0]
    dx, dw, db = np.empty_like(x), np.empty_like(w), np.empty((o,), x.dtype)
    getattr(lib(), "zo_linear_bwd" + s)(_p(x), _p(w), _p(dy), _p(dx), _p(dw), _p(db), I64(b), I64(i), I64(o))
    return dx, dw, db


def maxpool2d_fwd(x, kernel, stride, pad):
    x = _c(x)
    s, _ = _sfx(x)
    n, c, h, w = x.shape
    (kh, kw), (sh, sw), (ph, pw) = _pair(kernel), _pair(stride), _pair(pad)
    oh, ow = (h + 2 * ph - kh) // sh + 1, (w + 2 * pw - kw) // sw + 1
    y = np.empty((n, c, oh, ow), x.dtype)
    getattr(lib(), "zo_maxpool2d_fwd" + s)(_p(x), _p(y), *[I64(v) for v in (n, c, h, w, kh, kw, sh, sw, ph, pw)])
    return y


def maxpool2d_bwd(x, dy, kernel, stride, pad):
    x = _c(x)
    s, _ = _sfx(x)
    dy = _c(dy, x.dtype)
    n, c, h, w = x.shape
    (kh, kw), (sh, sw), (ph, pw) = _pair(kernel), _pair(stride), _pair(pad)
    dx = np.empty_like(x)
    getattr(lib(), "zo_maxpool2d_bwd" + s)(_p(x), _p(dy), _p(dx), *[I64(v) for v in (n, c, h, w, kh, kw, sh, sw, ph, pw)])
    return dx


def gap_fwd(x):
    x = _c(x)
    s, _ = _sfx(x)
    n, c, h, w = x.shape
    y = np.empty((n, c), x.dtype)
    getattr(lib(), "zo_gap_fwd" + s)(_p(x), _p(y), I64(n), I64(c), I64(h * w))
    return y


def gap_bwd(dy, hw_shape):
    dy = _c(dy)
    s, _ = _sfx(dy)
    n, c = dy.shape
    h, w = hw_shape
    dx = np.empty((n, c, h, w), dy.dtype)
    getattr(lib(), "zo_gap_bwd" + s)(_p(dy), _p(dx), I64(n), I64(c), I64(h * w))
    return dx


def softmax_xent(z, t, want_grad=True):
    """Returns (loss, dz)."""
    z = _c(z)
    s, ct = _sfx(z)
    t = _c(t, z.dtype)
    b, k = z.shape
    loss = ct(0)
    dz = np.empty_like(z) if want_grad else None
    getattr(lib(), "zo_softmax_xent" + s)(_p(z), _p(t), ctypes.byref(loss), _p(dz), I64(b), I64(k))
    return float(loss.value), dz


def sgd_step(p, g, lr):
    """In place on p."""
    s, ct = _sfx(p)
    assert p.flags.c_contiguous
    getattr(lib(), "zo_sgd_step" + s)(_p(p), _p(_c(g, p.dtype)), ct(lr), I64(p.size))


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step_t, weight_decay=0.0, decay=False):
    """In place on p, m, v.  step_t is the 1-based step count (adam.rs:21-22)."""
    s, ct = _sfx(p)
    assert p.flags.c_contiguous and m.flags.c_contiguous and v.flags.c_contiguous
    getattr(lib(), "zo_adam_step" + s)(_p(p), _p(_c(g, p.dtype)), _p(m), _p(v), ct(lr), ct(beta1), ct(beta2), ct(eps),
                                       ct(weight_decay), int(decay), I64(step_t), I64(p.size))


# ---- axis reductions / strided copies (numpy restatements; the accumulation ORDER is the reference's) ----------------------------
def sum_axis(a, axis, keep_dim=False):
    """Matrix::sum(axis, keep_dim), zenu-matrix/src/operation/sum.rs:9-31: result = zeros; for i in 0..shape[axis]: result += a[.., i, ..]
    (sequential accumulation in the element type, not numpy's pairwise sum)."""
    a = np.asarray(a)
    res = np.zeros(a.shape[:axis] + a.shape[axis + 1:], a.dtype)
    for i in range(a.shape[axis]):
        res += np.take(a, i, axis=axis)
    return np.expand_dims(res, axis) if keep_dim else res


def sum_to(a, shape):
    """sum_to(source, target), operation/sum.rs:35-92: leading extra axes are summed away one by one (sum(0)), then every axis whose
    target extent is 1 is summed with keep_dim."""
    a = np.asarray(a)
    shape = tuple(shape)
    assert a.ndim >= len(shape)
    while a.ndim > len(shape):
        a = sum_axis(a, 0)
    for k, (s, t) in enumerate(zip(a.shape, shape)):
        if s != t:
            assert t == 1, "sum_to: incompatible target shape"
            a = sum_axis(a, k, keep_dim=True)
    return a.copy()


def mean_axis(a, axis, keep_dim=False):
    """Matrix::mean(Some(axis), keep_dim), operation/mean.rs:8-20: sum / len in the element type."""
    a = np.asarray(a)
    return sum_axis(a, axis, keep_dim) / a.dtype.type(a.shape[axis])


def variance_axis(a, axis, keep_dim=False):
    """Matrix::variance(Some(axis), keep_dim), operation/var.rs:18-26: biased, mean((x - mean)^2)."""
    a = np.asarray(a)
    diff = a - mean_axis(a, axis, keep_dim=True)
    return mean_axis(diff * diff, axis, keep_dim)


def copy_strided(src_view):
    """copy_from into a default-stride destination (operation/copy_from.rs:57-123): element order of the source VIEW."""
    return np.ascontiguousarray(src_view)
