"""Host-side mirror of the reference's device-dispatched operator interface, bound to the C ABI.

Function names and argument meaning follow zenu-matrix's traits for this path
(`ConvFwd::conv_fwd` / `ConvBkwdData` / `ConvBkwdFilter` / `ConvBias`  nn/conv/interface.rs:37-89,
`BatchNormalization` nn/batch_norm.rs:130-167, `Gemm` operation/mul.rs:12-29, `ReluOps` operation/relu.rs:12-29,
`AddOps..DivOps` operation/basic_operations.rs:32-276).  Tensors are torch CUDA tensors used purely as device
memory; every op is one call into libzenu_b200.so on the Context's stream.  Shape errors raise
`ZenuB200Error` (the reference panics in shape_check, nn/conv/shape_check.rs:4-53).
"""
import ctypes
import weakref

import torch

from . import _lib
from ._lib import (ZB_F32, ZB_F64, ZB_MATH_DEFAULT, ZB_MATH_FP32, ZB_MATH_TF32, ZB_NCHW, ZB_NHWC, ConvDesc,
                   ZenuB200Error, check)

_DT = {torch.float32: ZB_F32, torch.float64: ZB_F64}
_OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3}


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


class Context:
    """One per process/rank (zb_ctx).  Shares torch's current CUDA stream so torch copies and our kernels order."""

    def __init__(self, device=None, math=ZB_MATH_TF32):
        if not torch.cuda.is_available():
            raise ZenuB200Error("zenu_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream().cuda_stream
        self._children = weakref.WeakSet()   # models created on this ctx: destroyed before the ctx (they use its streams)
        self._h = ctypes.c_void_p()
        check(self.lib.zb_ctx_create(ctypes.byref(self._h), self.device,
                                     ctypes.c_void_p(s if s != 0 else _lib.CUDA_STREAM_LEGACY)))
        check(self.lib.zb_ctx_set_math(self._h, math))

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            for child in list(self._children):
                child.close()
            self.lib.zb_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(self.lib.zb_ctx_synchronize(self._h))

    def check(self):
        check(self.lib.zb_ctx_check(self._h))

    def launch_count(self):
        return int(self.lib.zb_ctx_launch_count(self._h))

    def set_math(self, math):
        check(self.lib.zb_ctx_set_math(self._h, math))


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _chk(t, name):
    if t is None:
        return
    if not t.is_cuda or not t.is_contiguous() or t.dtype not in _DT:
        raise ZenuB200Error(f"{name}: expected a contiguous f32/f64 CUDA tensor (default stride, "
                            "zenu-matrix/src/nn/conv/interface.rs:270-281)")


def _chk_mask(t, name):
    if not t.is_cuda or not t.is_contiguous() or t.dtype != torch.int32:
        raise ZenuB200Error(f"{name}: expected a contiguous int32 CUDA tensor (1 bit per element)")


def _desc(x_shape, w_shape, layout, pad, stride, dil):
    if layout == ZB_NCHW:
        n, c, h, w = x_shape
        k, c2, kh, kw = w_shape
    else:
        n, h, w, c = x_shape
        k, kh, kw, c2 = w_shape
    if c != c2:
        raise ZenuB200Error(f"conv: input has {c} channels, filter expects {c2}")
    (ph, pw), (sh, sw), (dh, dw) = _pair(pad), _pair(stride), _pair(dil)
    return ConvDesc(n, c, h, w, k, kh, kw, ph, pw, sh, sw, dh, dw)


def conv_out_shape(x_shape, w_shape, layout, pad, stride, dil):
    d = _desc(x_shape, w_shape, layout, pad, stride, dil)
    lib = _lib.load()
    p = lib.zb_conv_out_size(d.h, d.kh, d.pad_h, d.stride_h, d.dil_h)
    q = lib.zb_conv_out_size(d.w, d.kw, d.pad_w, d.stride_w, d.dil_w)
    if p <= 0 or q <= 0 or d.h + 2 * d.pad_h < d.dil_h * (d.kh - 1) + 1 or d.w + 2 * d.pad_w < d.dil_w * (d.kw - 1) + 1:
        raise ZenuB200Error("conv: filter larger than the padded input (zenu-matrix/src/nn/conv/shape_check.rs:4-53)")
    return (d.n, d.k, p, q) if layout == ZB_NCHW else (d.n, p, q, d.k)


# ---- convolution (ConvFwd / ConvBkwdData / ConvBkwdFilter / ConvBias) --------------------------------------------
def conv_fwd(ctx, x, w, pad=0, stride=1, dil=1, bias=None, layout=ZB_NCHW, math=ZB_MATH_DEFAULT, out=None):
    _chk(x, "conv_fwd input"); _chk(w, "conv_fwd filter"); _chk(bias, "conv_fwd bias")
    d = _desc(tuple(x.shape), tuple(w.shape), layout, pad, stride, dil)
    y = out if out is not None else torch.empty(conv_out_shape(tuple(x.shape), tuple(w.shape), layout, pad, stride, dil),
                                                dtype=x.dtype, device=x.device)
    check(ctx.lib.zb_conv2d_fprop(ctx.handle, _DT[x.dtype], layout, math, ctypes.byref(d), _p(x), _p(w), _p(bias), _p(y)))
    return y


def conv_fwd_bnstats(ctx, x, w, shift, pad=0, stride=1, dil=1, bias=None, layout=ZB_NHWC, math=ZB_MATH_DEFAULT):
    """conv_fwd that also returns the BatchNorm statistics partials of its output, accumulated in the conv epilogue:
    (y, partial [rows][2][K], rows).  rows == 0: this shape / math mode cannot fuse them (y is complete either way)."""
    _chk(x, "conv_fwd input"); _chk(w, "conv_fwd filter"); _chk(bias, "conv_fwd bias"); _chk(shift, "conv_fwd shift")
    d = _desc(tuple(x.shape), tuple(w.shape), layout, pad, stride, dil)
    y = torch.empty(conv_out_shape(tuple(x.shape), tuple(w.shape), layout, pad, stride, dil), dtype=x.dtype, device=x.device)
    cap = ctx.lib.zb_conv2d_bnstats_rows(ctx.handle)
    partial = torch.zeros((cap, 2, d.k), dtype=x.dtype, device=x.device)
    rows = ctypes.c_int64(0)
    check(ctx.lib.zb_conv2d_fprop_bnstats(ctx.handle, _DT[x.dtype], layout, math, ctypes.byref(d), _p(x), _p(w), _p(bias), _p(y),
                                          _p(shift), _p(partial), ctypes.byref(rows)))
    return y, partial, int(rows.value)


def batch_norm_2d_forward_train_prestats(ctx, momentum, x, scale, bias, mean, variance, partial, rows, shift, layout=ZB_NHWC,
                                         residual=None, relu=False):
    """batch_norm_2d_forward_train whose statistics pass was done by conv_fwd_bnstats (x = that conv's output)."""
    n, c, h, w = _nkhw(x.shape, layout)
    y = torch.empty_like(x)
    sm = torch.empty((c,), dtype=x.dtype, device=x.device)
    si = torch.empty((c,), dtype=x.dtype, device=x.device)
    check(ctx.lib.zb_bn2d_fwd_train_prestats(ctx.handle, _DT[x.dtype], layout, n, c, h, w, float(momentum), _p(x), _p(scale), _p(bias),
                                             _p(mean), _p(variance), _p(sm), _p(si), _p(y), _p(residual), int(bool(relu)),
                                             _p(partial), int(rows), _p(shift)))
    return y, sm, si


def conv_bkwd_data(ctx, dy, w, x_shape, pad=0, stride=1, dil=1, layout=ZB_NCHW, math=ZB_MATH_DEFAULT):
    _chk(dy, "conv_bkwd_data dy"); _chk(w, "conv_bkwd_data filter")
    d = _desc(tuple(x_shape), tuple(w.shape), layout, pad, stride, dil)
    if tuple(dy.shape) != conv_out_shape(tuple(x_shape), tuple(w.shape), layout, pad, stride, dil):
        raise ZenuB200Error("conv_bkwd_data: dy shape does not match the conv geometry")
    dx = torch.empty(tuple(x_shape), dtype=dy.dtype, device=dy.device)
    check(ctx.lib.zb_conv2d_dgrad(ctx.handle, _DT[dy.dtype], layout, math, ctypes.byref(d), _p(dy), _p(w), _p(dx)))
    return dx


def conv_bkwd_data_accumulate(ctx, dy, w, dx, pad=0, stride=1, dil=1, layout=ZB_NCHW, math=ZB_MATH_DEFAULT):
    """dx += conv_bkwd_data(dy, w): the `grad + old` fan-in of Variable::set_grad (zenu-autograd/src/lib.rs:480-481)."""
    _chk(dy, "conv_bkwd_data dy"); _chk(w, "conv_bkwd_data filter"); _chk(dx, "conv_bkwd_data dx")
    d = _desc(tuple(dx.shape), tuple(w.shape), layout, pad, stride, dil)
    if tuple(dy.shape) != conv_out_shape(tuple(dx.shape), tuple(w.shape), layout, pad, stride, dil):
        raise ZenuB200Error("conv_bkwd_data: dy shape does not match the conv geometry")
    check(ctx.lib.zb_conv2d_dgrad_acc(ctx.handle, _DT[dy.dtype], layout, math, ctypes.byref(d), _p(dy), _p(w), _p(dx)))
    return dx


def conv_bkwd_data_accumulate_masked(ctx, dy, w, dx, mask, pad=0, stride=1, dil=1, layout=ZB_NCHW, math=ZB_MATH_DEFAULT):
    """dx = conv_bkwd_data(dy, w) + dx (.) mask, in place: the fan-in above with a lazily masked first arrival (bit e of `mask`,
    int32 words, keeps element e of dx in its memory order; the ReLU bits of batch_norm_2d_forward_train_masked)."""
    _chk(dy, "conv_bkwd_data dy"); _chk(w, "conv_bkwd_data filter"); _chk(dx, "conv_bkwd_data dx"); _chk_mask(mask, "conv_bkwd_data mask")
    d = _desc(tuple(dx.shape), tuple(w.shape), layout, pad, stride, dil)
    if tuple(dy.shape) != conv_out_shape(tuple(dx.shape), tuple(w.shape), layout, pad, stride, dil):
        raise ZenuB200Error("conv_bkwd_data: dy shape does not match the conv geometry")
    if mask.numel() * 32 < dx.numel():
        raise ZenuB200Error("conv_bkwd_data: mask shorter than dx")
    check(ctx.lib.zb_conv2d_dgrad_acc_masked(ctx.handle, _DT[dy.dtype], layout, math, ctypes.byref(d), _p(dy), _p(w), _p(dx), _p(mask)))
    return dx


def mask_apply(ctx, x, mask, out=None):
    """out[e] = x[e] if bit e of `mask` (int32 words) is set else 0."""
    _chk(x, "mask_apply x"); _chk_mask(mask, "mask_apply mask")
    if mask.numel() * 32 < x.numel():
        raise ZenuB200Error("mask_apply: mask shorter than x")
    out = torch.empty_like(x) if out is None else out
    check(ctx.lib.zb_mask_apply(ctx.handle, _DT[x.dtype], _p(x), _p(mask), _p(out), x.numel()))
    return out


def conv_bkwd_weight(ctx, dy, x, w_shape, pad=0, stride=1, dil=1, layout=ZB_NCHW, math=ZB_MATH_DEFAULT):
    _chk(dy, "conv_bkwd_weight dy"); _chk(x, "conv_bkwd_weight input")
    d = _desc(tuple(x.shape), tuple(w_shape), layout, pad, stride, dil)
    if tuple(dy.shape) != conv_out_shape(tuple(x.shape), tuple(w_shape), layout, pad, stride, dil):
        raise ZenuB200Error("conv_bkwd_weight: dy shape does not match the conv geometry")
    dw = torch.empty(tuple(w_shape), dtype=dy.dtype, device=dy.device)
    check(ctx.lib.zb_conv2d_wgrad(ctx.handle, _DT[dy.dtype], layout, math, ctypes.byref(d), _p(dy), _p(x), _p(dw)))
    return dw


PLAN_FPROP, PLAN_DGRAD, PLAN_WGRAD = 0, 1, 2
PLAN_BNSTATS, PLAN_BIAS, PLAN_ACCUMULATE = 1, 2, 4


def conv_plan_describe(ctx, op, x_shape, w_shape, pad=0, stride=1, dil=1, layout=ZB_NHWC, math=ZB_MATH_DEFAULT, flags=0,
                       dtype=torch.float32):
    """The launches zb_conv2d_{fprop,dgrad,wgrad} would make for this geometry (dry run of the same planners, nothing is
    launched or allocated): one "kernel<variant> key=value ...;" segment per launch; '~'-prefixed values depend on the batch."""
    d = _desc(tuple(x_shape), tuple(w_shape), layout, pad, stride, dil)
    need = ctx.lib.zb_conv2d_plan_describe(ctx.handle, op, _DT[dtype], layout, math, ctypes.byref(d), flags, None, 0)
    if need < 0:
        check(int(-need))
    buf = ctypes.create_string_buffer(int(need))
    ctx.lib.zb_conv2d_plan_describe(ctx.handle, op, _DT[dtype], layout, math, ctypes.byref(d), flags, buf, need)
    return buf.value.decode()


def plan_variant(text):
    """A plan description without its batch-size-dependent ('~') values."""
    return ";".join(" ".join(t for t in seg.split() if not t.startswith("~")) for seg in text.split(";") if seg.strip())


def plan_trace(ctx, enable=True):
    """Record the plan segments of the REAL conv / GEMM calls this thread makes from now on (read with plan_trace_read)."""
    check(ctx.lib.zb_ctx_plan_trace(ctx.handle, int(bool(enable))))


def plan_trace_read(ctx):
    need = ctx.lib.zb_ctx_plan_trace_read(ctx.handle, None, 0)
    buf = ctypes.create_string_buffer(int(need))
    ctx.lib.zb_ctx_plan_trace_read(ctx.handle, buf, need)
    return buf.value.decode()


def _nkhw(shape, layout):
    return (shape[0], shape[1], shape[2], shape[3]) if layout == ZB_NCHW else (shape[0], shape[3], shape[1], shape[2])


def conv2d_bias_add(ctx, x, bias, layout=ZB_NCHW):
    _chk(x, "conv2d_bias_add input"); _chk(bias, "bias")
    n, k, h, w = _nkhw(x.shape, layout)
    if bias.numel() != k:
        raise ZenuB200Error("conv2d_bias_add: bias must have C_out elements")
    y = torch.empty_like(x)
    check(ctx.lib.zb_conv2d_bias_add(ctx.handle, _DT[x.dtype], layout, _p(x), _p(bias), _p(y), n, k, h, w))
    return y


def conv2d_bias_bkwd(ctx, dy, layout=ZB_NCHW):
    _chk(dy, "conv2d_bias_bkwd dy")
    n, k, h, w = _nkhw(dy.shape, layout)
    db = torch.empty((k,), dtype=dy.dtype, device=dy.device)
    check(ctx.lib.zb_conv2d_bias_bwd(ctx.handle, _DT[dy.dtype], layout, _p(dy), _p(db), n, k, h, w))
    return db


# ---- batch norm (BatchNormalization) -------------------------------------------------------------------------------
def batch_norm_2d_forward_train(ctx, momentum, x, scale, bias, mean, variance, layout=ZB_NCHW, residual=None, relu=False):
    """Updates `mean` / `variance` (running stats) in place; returns (y, saving_mean, saving_inv_variance)."""
    for t, nm in ((x, "x"), (scale, "scale"), (bias, "bias"), (mean, "mean"), (variance, "variance"), (residual, "residual")):
        _chk(t, "batch_norm " + nm)
    n, c, h, w = _nkhw(x.shape, layout)
    if any(t.numel() != c for t in (scale, bias, mean, variance)):
        raise ZenuB200Error("batch_norm: scale/bias/mean/variance must have C elements (batch_norm.rs:424-487)")
    y = torch.empty_like(x)
    sm = torch.empty((c,), dtype=x.dtype, device=x.device)
    si = torch.empty((c,), dtype=x.dtype, device=x.device)
    check(ctx.lib.zb_bn2d_fwd_train(ctx.handle, _DT[x.dtype], layout, n, c, h, w, float(momentum), _p(x), _p(scale), _p(bias),
                                    _p(mean), _p(variance), _p(sm), _p(si), _p(y), _p(residual), int(bool(relu))))
    return y, sm, si


def batch_norm_2d_forward_train_masked(ctx, momentum, x, scale, bias, mean, variance, residual=None):
    """NHWC f32 fused BN(+residual)+ReLU forward that also writes the 1-bit ReLU mask: (y, saving_mean, saving_inv, mask)."""
    n, c, h, w = _nkhw(x.shape, ZB_NHWC)
    y = torch.empty_like(x)
    sm = torch.empty((c,), dtype=x.dtype, device=x.device)
    si = torch.empty((c,), dtype=x.dtype, device=x.device)
    mask = torch.zeros((int(ctx.lib.zb_bn2d_relu_mask_words(n, c, h, w)),), dtype=torch.int32, device=x.device)
    check(ctx.lib.zb_bn2d_fwd_train_fused(ctx.handle, _DT[x.dtype], ZB_NHWC, n, c, h, w, float(momentum), _p(x), _p(scale), _p(bias),
                                          _p(mean), _p(variance), _p(sm), _p(si), _p(y), _p(residual), 1, None, 0, None, _p(mask)))
    return y, sm, si, mask


def batch_norm_2d_backward_masked(ctx, x, y_grad, scale, saving_mean, saving_inv_variance, mask, want_residual_grad=True):
    """Backward of the fused BN+add+ReLU with the ReLU mask taken from `mask`: (x_grad, scale_grad, bias_grad, residual_grad).
    want_residual_grad=False: the masked gradient y_grad (.) mask is not written out (residual_grad is None; the caller keeps
    (y_grad, mask) and masks where it is consumed)."""
    n, c, h, w = _nkhw(x.shape, ZB_NHWC)
    dx, dres = torch.empty_like(x), (torch.empty_like(x) if want_residual_grad else None)
    ds = torch.empty((c,), dtype=x.dtype, device=x.device)
    db = torch.empty((c,), dtype=x.dtype, device=x.device)
    check(ctx.lib.zb_bn2d_bwd_mask(ctx.handle, _DT[x.dtype], ZB_NHWC, n, c, h, w, _p(x), _p(y_grad), _p(scale), _p(saving_mean),
                                   _p(saving_inv_variance), _p(dx), _p(ds), _p(db), _p(mask), _p(dres)))
    return dx, ds, db, dres


def batch_norm_2d_backward(ctx, x, y_grad, scale, saving_mean=None, saving_inv_variance=None, layout=ZB_NCHW,
                           y=None, want_residual_grad=False):
    """Returns (x_grad, scale_grad, bias_grad[, residual_grad]).  Pass the fused forward output in `y` when it had relu."""
    for t, nm in ((x, "x"), (y_grad, "y_grad"), (scale, "scale"), (saving_mean, "saving_mean"),
                  (saving_inv_variance, "saving_inv_variance"), (y, "y")):
        _chk(t, "batch_norm_backward " + nm)
    n, c, h, w = _nkhw(x.shape, layout)
    dx = torch.empty_like(x)
    ds = torch.empty((c,), dtype=x.dtype, device=x.device)
    db = torch.empty((c,), dtype=x.dtype, device=x.device)
    dres = torch.empty_like(x) if want_residual_grad else None
    check(ctx.lib.zb_bn2d_bwd(ctx.handle, _DT[x.dtype], layout, n, c, h, w, _p(x), _p(y_grad), _p(scale), _p(saving_mean),
                              _p(saving_inv_variance), _p(dx), _p(ds), _p(db), _p(y), _p(dres)))
    return (dx, ds, db, dres) if want_residual_grad else (dx, ds, db)


def batch_norm_2d_relu_backward(ctx, x, y_grad, scale, bias, saving_mean, saving_inv_variance, layout=ZB_NCHW):
    """Backward of relu(batch_norm(x)) (no residual) that recomputes the ReLU mask from x instead of reading y."""
    for t, nm in ((x, "x"), (y_grad, "y_grad"), (scale, "scale"), (bias, "bias"), (saving_mean, "saving_mean"),
                  (saving_inv_variance, "saving_inv_variance")):
        _chk(t, "batch_norm_relu_backward " + nm)
    n, c, h, w = _nkhw(x.shape, layout)
    dx = torch.empty_like(x)
    ds = torch.empty((c,), dtype=x.dtype, device=x.device)
    db = torch.empty((c,), dtype=x.dtype, device=x.device)
    check(ctx.lib.zb_bn2d_relu_bwd(ctx.handle, _DT[x.dtype], layout, n, c, h, w, _p(x), _p(y_grad), _p(scale), _p(bias),
                                   _p(saving_mean), _p(saving_inv_variance), _p(dx), _p(ds), _p(db)))
    return dx, ds, db


def batch_norm_2d_forward_inference(ctx, x, scale, bias, mean, variance, layout=ZB_NCHW):
    for t, nm in ((x, "x"), (scale, "scale"), (bias, "bias"), (mean, "mean"), (variance, "variance")):
        _chk(t, "batch_norm_inference " + nm)
    n, c, h, w = _nkhw(x.shape, layout)
    y = torch.empty_like(x)
    check(ctx.lib.zb_bn2d_fwd_infer(ctx.handle, _DT[x.dtype], layout, n, c, h, w, _p(x), _p(scale), _p(bias), _p(mean),
                                    _p(variance), _p(y)))
    return y


# ---- GEMM / Linear ---------------------------------------------------------------------------------------------------
def gemm(ctx, a, b, trans_a=False, trans_b=False, alpha=1.0, beta=0.0, c=None, math=ZB_MATH_DEFAULT):
    """Row-major C = alpha*op(A)*op(B) + beta*C (Gemm::gemm_unchecked)."""
    _chk(a, "gemm a"); _chk(b, "gemm b"); _chk(c, "gemm c")
    m, k = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    k2, n = (b.shape[1], b.shape[0]) if trans_b else (b.shape[0], b.shape[1])
    if k != k2:
        raise ZenuB200Error(f"gemm: inner dimensions differ ({k} vs {k2})")
    if c is None:
        c = torch.zeros((m, n), dtype=a.dtype, device=a.device) if beta != 0.0 else torch.empty((m, n), dtype=a.dtype, device=a.device)
    check(ctx.lib.zb_gemm(ctx.handle, _DT[a.dtype], math, int(trans_a), int(trans_b), m, n, k, float(alpha), _p(a), a.shape[1],
                          _p(b), b.shape[1], float(beta), _p(c), c.shape[1]))
    return c


def matmul(ctx, a, b, math=ZB_MATH_DEFAULT):
    return gemm(ctx, a, b, math=math)


def linear_fwd(ctx, x, w, bias=None, math=ZB_MATH_DEFAULT):
    _chk(x, "linear x"); _chk(w, "linear w"); _chk(bias, "linear bias")
    b, i = x.shape
    o, i2 = w.shape
    if i != i2:
        raise ZenuB200Error("linear: in_features mismatch")
    y = torch.empty((b, o), dtype=x.dtype, device=x.device)
    check(ctx.lib.zb_linear_fwd(ctx.handle, _DT[x.dtype], math, _p(x), _p(w), _p(bias), _p(y), b, i, o))
    return y


def linear_bwd(ctx, x, w, dy, math=ZB_MATH_DEFAULT):
    b, i = x.shape
    o = w.shape[0]
    dx, dw = torch.empty_like(x), torch.empty_like(w)
    db = torch.empty((o,), dtype=x.dtype, device=x.device)
    check(ctx.lib.zb_linear_bwd(ctx.handle, _DT[x.dtype], math, _p(x), _p(w), _p(dy), _p(dx), _p(dw), _p(db), b, i, o))
    return dx, dw, db


# ---- elementwise -------------------------------------------------------------------------------------------------------
def relu(ctx, x, alpha=0.0):
    _chk(x, "relu x")
    y = torch.empty_like(x)
    check(ctx.lib.zb_relu(ctx.handle, _DT[x.dtype], _p(x), _p(y), float(alpha), x.numel()))
    return y


def relu_backward_mask(ctx, x, alpha=0.0):
    _chk(x, "relu_backward_mask x")
    y = torch.empty_like(x)
    check(ctx.lib.zb_relu_backward_mask(ctx.handle, _DT[x.dtype], _p(x), _p(y), float(alpha), x.numel()))
    return y


def relu_bwd(ctx, x, dy, alpha=0.0):
    dx = torch.empty_like(x)
    check(ctx.lib.zb_relu_bwd(ctx.handle, _DT[x.dtype], _p(x), _p(dy), _p(dx), float(alpha), x.numel()))
    return dx


def binary(ctx, op, a, b, out=None):
    """a op b; b may be a same-shape tensor, a python scalar, or a 1-D tensor broadcast over the last axis."""
    _chk(a, "binary a")
    out = torch.empty_like(a) if out is None else out
    if isinstance(b, (int, float)):
        check(ctx.lib.zb_binary_scalar(ctx.handle, _DT[a.dtype], _OPS[op], _p(a), float(b), _p(out), a.numel()))
    elif b.shape == a.shape:
        check(ctx.lib.zb_binary(ctx.handle, _DT[a.dtype], _OPS[op], _p(a), _p(b), _p(out), a.numel()))
    elif b.dim() == 1 and b.shape[0] == a.shape[-1]:
        check(ctx.lib.zb_binary_bcast_rows(ctx.handle, _DT[a.dtype], _OPS[op], _p(a), _p(b), _p(out), a.numel() // a.shape[-1], a.shape[-1]))
    else:
        raise ZenuB200Error("binary: unsupported broadcast")
    return out


def add(ctx, a, b, out=None):
    return binary(ctx, "add", a, b, out)


def sum_rows(ctx, a):
    _chk(a, "sum_rows a")
    out = torch.empty((a.shape[-1],), dtype=a.dtype, device=a.device)
    check(ctx.lib.zb_sum_rows(ctx.handle, _DT[a.dtype], _p(a), _p(out), a.numel() // a.shape[-1], a.shape[-1]))
    return out


def _axis_view(shape, axis):
    outer = 1
    for v in shape[:axis]:
        outer *= v
    inner = 1
    for v in shape[axis + 1:]:
        inner *= v
    return outer, shape[axis], inner


def _reduced_shape(shape, axis, keep_dim):
    return tuple(shape[:axis]) + ((1,) if keep_dim else ()) + tuple(shape[axis + 1:])


def sum_axis(ctx, a, axis, keep_dim=False):
    """Matrix::sum(axis, keep_dim) (zenu-matrix/src/operation/sum.rs:9-31)."""
    _chk(a, "sum input")
    outer, n, inner = _axis_view(tuple(a.shape), axis)
    out = torch.empty(_reduced_shape(tuple(a.shape), axis, keep_dim), dtype=a.dtype, device=a.device)
    check(ctx.lib.zb_sum_axis(ctx.handle, _DT[a.dtype], _p(a), _p(out), outer, n, inner))
    return out


def mean_axis(ctx, a, axis, keep_dim=False):
    """Matrix::mean(Some(axis), keep_dim) (operation/mean.rs:8-20)."""
    _chk(a, "mean input")
    outer, n, inner = _axis_view(tuple(a.shape), axis)
    out = torch.empty(_reduced_shape(tuple(a.shape), axis, keep_dim), dtype=a.dtype, device=a.device)
    check(ctx.lib.zb_mean_axis(ctx.handle, _DT[a.dtype], _p(a), _p(out), outer, n, inner))
    return out


def variance_axis(ctx, a, axis, keep_dim=False, want_mean=False):
    """Matrix::variance(Some(axis), keep_dim) (operation/var.rs:18-26): biased."""
    _chk(a, "variance input")
    outer, n, inner = _axis_view(tuple(a.shape), axis)
    shp = _reduced_shape(tuple(a.shape), axis, keep_dim)
    out = torch.empty(shp, dtype=a.dtype, device=a.device)
    mean = torch.empty(shp, dtype=a.dtype, device=a.device) if want_mean else None
    check(ctx.lib.zb_variance_axis(ctx.handle, _DT[a.dtype], _p(a), _p(out), _p(mean), outer, n, inner))
    return (out, mean) if want_mean else out


def sum_to(ctx, a, shape):
    """sum_to(source, target) (operation/sum.rs:35-92): reduce `a` to a shape it broadcasts from (right-aligned)."""
    _chk(a, "sum_to input")
    shape = tuple(int(v) for v in shape)
    out = torch.empty(shape, dtype=a.dtype, device=a.device)
    ss = (ctypes.c_int64 * max(a.dim(), 1))(*a.shape)
    ds = (ctypes.c_int64 * max(len(shape), 1))(*shape)
    check(ctx.lib.zb_sum_to(ctx.handle, _DT[a.dtype], _p(a), ss, a.dim(), _p(out), ds, len(shape)))
    return out


def fill(ctx, x, value):
    """DeviceBase::zeros and friends: x[:] = value (the reference zero-fills by scaling with 0, NaN-unsafe)."""
    _chk(x, "fill target")
    check(ctx.lib.zb_fill(ctx.handle, _DT[x.dtype], _p(x), float(value), x.numel()))
    return x


def copy(ctx, src, dst=None):
    _chk(src, "copy source")
    dst = torch.empty_like(src) if dst is None else dst
    check(ctx.lib.zb_copy(ctx.handle, _DT[src.dtype], _p(src), _p(dst), src.numel()))
    return dst


def copy_strided(ctx, src, dst):
    """copy_from for arbitrarily strided views (CopyBlas::copy_raw, operation/copy_from.rs:9-55): src and dst are torch views of the
    same shape with any non-negative element strides (transposes, slices with steps, broadcast sources)."""
    if tuple(src.shape) != tuple(dst.shape) or src.dtype != dst.dtype or src.dtype not in _DT:
        raise ZenuB200Error("copy_strided: shape / dtype mismatch")
    nd = src.dim()
    mk = lambda v: (ctypes.c_int64 * max(nd, 1))(*v)  # noqa: E731
    check(ctx.lib.zb_copy_strided(ctx.handle, _DT[src.dtype], ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()), nd,
                                  mk(src.shape), mk(src.stride()), mk(dst.stride())))
    return dst


def to_nhwc(ctx, x):
    n, c, h, w = x.shape
    y = torch.empty((n, h, w, c), dtype=x.dtype, device=x.device)
    check(ctx.lib.zb_nchw_to_nhwc(ctx.handle, _DT[x.dtype], _p(x), _p(y), n, c, h, w))
    return y


def to_nchw(ctx, x):
    n, h, w, c = x.shape
    y = torch.empty((n, c, h, w), dtype=x.dtype, device=x.device)
    check(ctx.lib.zb_nhwc_to_nchw(ctx.handle, _DT[x.dtype], _p(x), _p(y), n, c, h, w))
    return y


# ---- pooling / loss ("next" rows) ----------------------------------------------------------------------------------------
def max_pool_2d(ctx, x, kernel, stride, pad, layout=ZB_NCHW):
    n, c, h, w = _nkhw(x.shape, layout)
    (kh, kw), (sh, sw), (ph, pw) = _pair(kernel), _pair(stride), _pair(pad)
    p, q = (h + 2 * ph - kh) // sh + 1, (w + 2 * pw - kw) // sw + 1
    y = torch.empty((n, c, p, q) if layout == ZB_NCHW else (n, p, q, c), dtype=x.dtype, device=x.device)
    check(ctx.lib.zb_maxpool2d_fwd(ctx.handle, _DT[x.dtype], layout, _p(x), _p(y), n, c, h, w, kh, kw, sh, sw, ph, pw))
    return y


def max_pool_2d_backward(ctx, x, dy, kernel, stride, pad, layout=ZB_NCHW):
    n, c, h, w = _nkhw(x.shape, layout)
    (kh, kw), (sh, sw), (ph, pw) = _pair(kernel), _pair(stride), _pair(pad)
    dx = torch.empty_like(x)
    check(ctx.lib.zb_maxpool2d_bwd(ctx.handle, _DT[x.dtype], layout, _p(x), _p(dy), _p(dx), n, c, h, w, kh, kw, sh, sw, ph, pw))
    return dx


def max_pool_2d_indexed(ctx, x, kernel, stride, pad):
    """NHWC only.  Returns (y, idx): idx[N,P,Q,C] uint8 = winning tap r*kw+s (255 = a padding zero won)."""
    n, c, h, w = _nkhw(x.shape, ZB_NHWC)
    (kh, kw), (sh, sw), (ph, pw) = _pair(kernel), _pair(stride), _pair(pad)
    p, q = (h + 2 * ph - kh) // sh + 1, (w + 2 * pw - kw) // sw + 1
    y = torch.empty((n, p, q, c), dtype=x.dtype, device=x.device)
    idx = torch.empty((n, p, q, c), dtype=torch.uint8, device=x.device)
    check(ctx.lib.zb_maxpool2d_fwd_idx(ctx.handle, _DT[x.dtype], ZB_NHWC, _p(x), _p(y), ctypes.c_void_p(idx.data_ptr()), n, c, h, w,
                                       kh, kw, sh, sw, ph, pw))
    return y, idx


def max_pool_2d_indexed_backward(ctx, dy, idx, x_shape, kernel, stride, pad):
    n, c, h, w = _nkhw(x_shape, ZB_NHWC)
    (kh, kw), (sh, sw), (ph, pw) = _pair(kernel), _pair(stride), _pair(pad)
    dx = torch.empty(tuple(x_shape), dtype=dy.dtype, device=dy.device)
    check(ctx.lib.zb_maxpool2d_bwd_idx(ctx.handle, _DT[dy.dtype], ZB_NHWC, _p(dy), ctypes.c_void_p(idx.data_ptr()), _p(dx), n, c, h, w,
                                       kh, kw, sh, sw, ph, pw))
    return dx


def batch_norm_relu_max_pool_forward_train(ctx, momentum, x, scale, bias, mean, variance, stat_partial=None, stat_rows=0, shift=None,
                                           kernel=3, stride=2, pad=1):
    """NHWC f32: BatchNorm2d(train) + ReLU + max_pool_2d(3, 2, 1) in one pass (the ResNet stem).  Returns (y_pool, idx, saving_mean,
    saving_inv_variance); running mean / variance are updated in place.  Raises ZenuB200Error (unsupported) for other geometries."""
    for t, nm in ((x, "x"), (scale, "scale"), (bias, "bias"), (mean, "mean"), (variance, "variance")):
        _chk(t, "batch_norm_relu_max_pool " + nm)
    n, c, h, w = _nkhw(x.shape, ZB_NHWC)
    p, q = (h + 2 * pad - kernel) // stride + 1, (w + 2 * pad - kernel) // stride + 1
    y = torch.empty((n, p, q, c), dtype=x.dtype, device=x.device)
    idx = torch.empty((n, p, q, c), dtype=torch.uint8, device=x.device)
    sm = torch.empty((c,), dtype=x.dtype, device=x.device)
    si = torch.empty((c,), dtype=x.dtype, device=x.device)
    check(ctx.lib.zb_bn2d_relu_maxpool_fwd_train(ctx.handle, _DT[x.dtype], ZB_NHWC, n, c, h, w, kernel, stride, pad, float(momentum), _p(x),
                                                 _p(scale), _p(bias), _p(mean), _p(variance), _p(sm), _p(si), _p(y),
                                                 ctypes.c_void_p(idx.data_ptr()), _p(stat_partial), int(stat_rows), _p(shift)))
    return y, idx, sm, si


def batch_norm_relu_max_pool_backward(ctx, x, dy_pool, idx, scale, bias, saving_mean, saving_inv_variance, kernel=3, stride=2, pad=1):
    """Backward of batch_norm_relu_max_pool_forward_train: (x_grad, scale_grad, bias_grad)."""
    n, c, h, w = _nkhw(x.shape, ZB_NHWC)
    dx = torch.empty_like(x)
    ds = torch.empty((c,), dtype=x.dtype, device=x.device)
    db = torch.empty((c,), dtype=x.dtype, device=x.device)
    check(ctx.lib.zb_bn2d_relu_maxpool_bwd(ctx.handle, _DT[x.dtype], ZB_NHWC, n, c, h, w, kernel, stride, pad, _p(x), _p(dy_pool),
                                           ctypes.c_void_p(idx.data_ptr()), _p(scale), _p(bias), _p(saving_mean), _p(saving_inv_variance),
                                           _p(dx), _p(ds), _p(db)))
    return dx, ds, db


def global_avg_pool(ctx, x, layout=ZB_NCHW):
    n, c, h, w = _nkhw(x.shape, layout)
    y = torch.empty((n, c), dtype=x.dtype, device=x.device)
    check(ctx.lib.zb_gap_fwd(ctx.handle, _DT[x.dtype], layout, _p(x), _p(y), n, c, h * w))
    return y


def global_avg_pool_backward(ctx, dy, x_shape, layout=ZB_NCHW):
    n, c, h, w = _nkhw(x_shape, layout)
    dx = torch.empty(tuple(x_shape), dtype=dy.dtype, device=dy.device)
    check(ctx.lib.zb_gap_bwd(ctx.handle, _DT[dy.dtype], layout, _p(dy), _p(dx), n, c, h * w))
    return dx


def softmax_cross_entropy(ctx, z, t, want_grad=True):
    loss = torch.empty((1,), dtype=z.dtype, device=z.device)
    dz = torch.empty_like(z) if want_grad else None
    check(ctx.lib.zb_softmax_xent(ctx.handle, _DT[z.dtype], _p(z), _p(t), _p(loss), _p(dz), z.shape[0], z.shape[1]))
    return loss, dz


# ---- optimizers ------------------------------------------------------------------------------------------------------------
def sgd_step(ctx, param, grad, lr, grad_scale=1.0):
    check(ctx.lib.zb_sgd_step(ctx.handle, _DT[param.dtype], _p(param), _p(grad), float(lr), float(grad_scale), param.numel()))


def adam_step(ctx, param, grad, m, v, lr, beta1, beta2, eps, step_t, weight_decay=0.0, decay=False, grad_scale=1.0):
    check(ctx.lib.zb_adam_step(ctx.handle, _DT[param.dtype], _p(param), _p(grad), _p(m), _p(v), float(lr), float(beta1),
                               float(beta2), float(eps), float(weight_decay), int(bool(decay)), int(step_t), float(grad_scale),
                               param.numel()))


# ---- input pipeline ---------------------------------------------------------------------------------------------------
def input_u8_to_float(ctx, src_u8, mean=None, std=None, src_layout=ZB_NHWC, dtype=torch.float32):
    """uint8 batch ([N,H,W,C] for ZB_NHWC, [N,C,H,W] for ZB_NCHW) -> normalised NCHW float batch on the device."""
    if not src_u8.is_cuda or not src_u8.is_contiguous() or src_u8.dtype != torch.uint8:
        raise ZenuB200Error("input_u8_to_float: contiguous uint8 CUDA tensor expected")
    if src_layout == ZB_NHWC:
        n, h, w, c = src_u8.shape
    else:
        n, c, h, w = src_u8.shape
    out = torch.empty((n, c, h, w), dtype=dtype, device=src_u8.device)
    m = (ctypes.c_double * c)(*[float(v) for v in mean]) if mean is not None else None
    s = (ctypes.c_double * c)(*[float(v) for v in std]) if std is not None else None
    check(ctx.lib.zb_input_u8_to_float(ctx.handle, _DT[dtype], src_layout, _p(src_u8), _p(out), n, c, h, w, m, s))
    return out


def onehot(ctx, labels, classes, dtype=torch.float32):
    if not labels.is_cuda or not labels.is_contiguous() or labels.dtype != torch.int32:
        raise ZenuB200Error("onehot: contiguous int32 CUDA labels expected")
    out = torch.empty((labels.numel(), classes), dtype=dtype, device=labels.device)
    check(ctx.lib.zb_onehot(ctx.handle, _DT[dtype], _p(labels), _p(out), labels.numel(), classes))
    return out
