"""Functional tests of the two compatibility surfaces, called exactly as the reference's bindgen'd Rust would call them
(raw device pointers, int sizes / strides, the prototypes of tests/golden/ref_symbols.json):

  * every one of the 126 zenu-cuda-kernel-sys symbols (zenu-cuda-kernel-sys/kernel/*.h), unit-stride AND strided, against numpy /
    the CPU oracle — including the cases where the reference kernels are wrong and this library deliberately is not
    (conv2d_bias_bkwd with N > 1, array_clip's swapped indices, array_max_idx's host/device mix-up);
  * the cuDNN-frontend wrapper's conv backward-data / backward-filter descriptors (cudnn_frontend_wrapper.h:132-186) and its
    BatchNorm forward / backward descriptors (:33-98) with cuDNN-frontend's conventions (momentum on the NEW statistic).
"""
import ctypes
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import zenu_oracle as zo  # noqa: E402

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CT = {"float": ctypes.c_float, "double": ctypes.c_double, "int": ctypes.c_int}


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from zenu_b200 import _lib
    lib = _lib.load()
    with open(os.path.join(ROOT, "tests", "golden", "ref_symbols.json")) as f:
        ref = json.load(f)["kernel_sys"]
    for name, proto in ref.items():   # argtypes straight from the REFERENCE prototypes
        fn = getattr(lib, name)
        fn.restype = None
        fn.argtypes = [ctypes.c_void_p if a.endswith("*") else CT[a] for a in proto["args"]]
    return lib, ref


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.detach().cpu().numpy()


def np_t(sfx):
    return np.float32 if sfx == "float" else np.float64


N = 1000
STRIDES = [(1, 1), (3, 2), (2, 1)]   # (input stride, output stride)


@pytest.mark.parametrize("sfx", ["float", "double"])
def test_kernel_sys_binary_and_scalar_families(lib, sfx):
    lib, _ = lib
    T = np_t(sfx)
    rng = np.random.default_rng(1)
    ops = {"add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide}
    for sa, so in STRIDES:
        a = rng.standard_normal(N * sa).astype(T)
        b = (rng.standard_normal(N * 2) + 3.0).astype(T)
        for op, f in ops.items():
            A, B, C = dev(a), dev(b), dev(np.zeros(N * so, T))
            getattr(lib, f"array_array_{op}_{sfx}")(A.data_ptr(), sa, B.data_ptr(), 2, C.data_ptr(), so, N)
            np.testing.assert_array_equal(host(C)[::so][:N], f(a[::sa][:N], b[::2][:N]))
            if so > 1:
                assert not host(C)[1::so].any()          # gaps of a strided output stay untouched
            A2 = dev(a)
            getattr(lib, f"array_array_{op}_assign_{sfx}")(A2.data_ptr(), sa, B.data_ptr(), 2, N)
            np.testing.assert_array_equal(host(A2)[::sa][:N], f(a[::sa][:N], b[::2][:N]))
            s = T(1.75)
            O = dev(np.zeros(N * so, T))
            getattr(lib, f"array_scalar_{op}_{sfx}")(A.data_ptr(), N, sa, s, O.data_ptr(), so)
            np.testing.assert_allclose(host(O)[::so][:N], f(a[::sa][:N], s), rtol=1e-6)
            A3 = dev(a)
            getattr(lib, f"array_scalar_{op}_assign_{sfx}")(A3.data_ptr(), N, sa, s)
            np.testing.assert_allclose(host(A3)[::sa][:N], f(a[::sa][:N], s), rtol=1e-6)
            S = dev(np.array([2.5], T))                   # scalar read from device memory
            O2 = dev(np.zeros(N * so, T))
            getattr(lib, f"array_scalar_pointer_{op}_{sfx}")(A.data_ptr(), N, sa, S.data_ptr(), O2.data_ptr(), so)
            np.testing.assert_array_equal(host(O2)[::so][:N], f(a[::sa][:N], T(2.5)))
            A4 = dev(a)
            getattr(lib, f"array_scalar_pointer_{op}_assign_{sfx}")(A4.data_ptr(), N, sa, S.data_ptr())
            np.testing.assert_array_equal(host(A4)[::sa][:N], f(a[::sa][:N], T(2.5)))


@pytest.mark.parametrize("sfx", ["float", "double"])
def test_kernel_sys_unary_clip_pow_relu(lib, sfx):
    lib, _ = lib
    T = np_t(sfx)
    rng = np.random.default_rng(2)
    tol = 2e-6 if sfx == "float" else 1e-12
    fam = {"sin": (np.sin, "n"), "cos": (np.cos, "n"), "tan": (np.tan, "u"), "asin": (np.arcsin, "u"), "acos": (np.arccos, "u"),
           "atan": (np.arctan, "n"), "sinh": (np.sinh, "n"), "cosh": (np.cosh, "n"), "tanh": (np.tanh, "n"), "abs": (np.abs, "n"),
           "sqrt": (np.sqrt, "p"), "exp": (np.exp, "n"), "log": (np.log, "p")}
    for sa, so in STRIDES:
        for name, (f, dom) in fam.items():
            a = {"n": rng.standard_normal(N * sa), "u": rng.uniform(-0.9, 0.9, N * sa), "p": rng.uniform(0.1, 4.0, N * sa)}[dom].astype(T)
            A, O = dev(a), dev(np.zeros(N * so, T))
            getattr(lib, f"array_{name}_{sfx}")(A.data_ptr(), N, sa, O.data_ptr(), so)
            np.testing.assert_allclose(host(O)[::so][:N], f(a[::sa][:N]), rtol=tol, atol=tol)
            A2 = dev(a)
            getattr(lib, f"array_{name}_assign_{sfx}")(A2.data_ptr(), N, sa)
            np.testing.assert_allclose(host(A2)[::sa][:N], f(a[::sa][:N]), rtol=tol, atol=tol)
        a = rng.standard_normal(N * sa).astype(T)
        A, O = dev(a), dev(np.zeros(N * so, T))
        # clip: the reference indexes the input with the OUTPUT stride and vice versa (array_scalar.cu:9): not reproduced
        getattr(lib, f"array_clip_{sfx}")(A.data_ptr(), O.data_ptr(), N, sa, so, T(-0.5), T(0.25))
        np.testing.assert_array_equal(host(O)[::so][:N], np.clip(a[::sa][:N], -0.5, 0.25))
        A2 = dev(a)
        getattr(lib, f"array_clip_assign_{sfx}")(A2.data_ptr(), N, sa, T(-0.5), T(0.25))
        np.testing.assert_array_equal(host(A2)[::sa][:N], np.clip(a[::sa][:N], -0.5, 0.25))
        M = dev(np.zeros(N * so, T))
        getattr(lib, f"array_clip_backward_{sfx}")(A.data_ptr(), M.data_ptr(), T(0.25), T(-0.5), N, sa, so)   # (max, min) order
        inside = ((a[::sa][:N] >= -0.5) & (a[::sa][:N] <= 0.25)).astype(T)
        np.testing.assert_array_equal(host(M)[::so][:N], inside)
        A3 = dev(a)
        getattr(lib, f"array_clip_backward_assign_{sfx}")(A3.data_ptr(), T(0.25), T(-0.5), N, sa)
        np.testing.assert_array_equal(host(A3)[::sa][:N], inside)
        p = np.abs(a) + T(0.5)
        P, O2 = dev(p), dev(np.zeros(N * so, T))
        getattr(lib, f"array_pow_{sfx}")(P.data_ptr(), N, sa, T(1.5), O2.data_ptr(), so)
        np.testing.assert_allclose(host(O2)[::so][:N], p[::sa][:N] ** T(1.5), rtol=10 * tol)
        P2 = dev(p)
        getattr(lib, f"array_pow_assign_{sfx}")(P2.data_ptr(), N, sa, T(1.5))
        np.testing.assert_allclose(host(P2)[::sa][:N], p[::sa][:N] ** T(1.5), rtol=10 * tol)
        # relu / mask with a leaky slope (zenu-matrix/src/operation/relu.rs:31-91: mask = x > 0 ? 1 : -alpha ... as the kernel has it)
        R = dev(np.zeros(N * so, T))
        getattr(lib, f"relu_{sfx}")(A.data_ptr(), R.data_ptr(), T(0.1), N, sa, so)
        np.testing.assert_array_equal(host(R)[::so][:N], zo.relu(a[::sa][:N].copy(), 0.1))
        getattr(lib, f"relu_backward_mask_{sfx}")(A.data_ptr(), R.data_ptr(), T(0.1), N, sa, so)
        np.testing.assert_array_equal(host(R)[::so][:N], zo.relu_backward_mask(a[::sa][:N].copy(), 0.1))


@pytest.mark.parametrize("sfx", ["float", "double"])
def test_kernel_sys_conv_bias_memory_argmax(lib, sfx):
    lib, _ = lib
    T = np_t(sfx)
    rng = np.random.default_rng(3)
    n, c, h, w = 3, 5, 4, 6
    x = rng.standard_normal((n, c, h, w)).astype(T)
    b = rng.standard_normal(c).astype(T)
    X, B, Y = dev(x), dev(b), dev(np.zeros_like(x))
    getattr(lib, f"conv_bias_add_{sfx}")(X.data_ptr(), Y.data_ptr(), h * w, B.data_ptr(), c, x.size)
    np.testing.assert_array_equal(host(Y), zo.conv2d_bias_add(x, b))
    DB = dev(np.zeros(c, T))
    getattr(lib, f"conv2d_bias_bkwd_{sfx}")(X.data_ptr(), DB.data_ptr(), n, c, h, w)     # N = 3: the reference kernel is wrong here (S4)
    np.testing.assert_allclose(host(DB).ravel(), np.asarray(zo.conv2d_bias_bkwd(x)).ravel(), rtol=1e-5)
    a = rng.standard_normal(64).astype(T)
    A = dev(a)
    out = CT[sfx](0)
    getattr(lib, f"memory_access_{sfx}").argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(CT[sfx])]
    getattr(lib, f"memory_access_{sfx}")(A.data_ptr(), 17, ctypes.byref(out))
    assert out.value == a[17]
    getattr(lib, f"memory_set_{sfx}")(A.data_ptr(), 17, T(42.5))
    assert host(A)[17] == T(42.5) and host(A)[16] == a[16]
    for size, stride in ((64, 1), (5000, 1), (1000, 3)):
        v = rng.standard_normal(size * stride).astype(T)
        V = dev(v)
        idx = ctypes.c_int(-1)
        getattr(lib, f"array_max_idx_{sfx}").argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        getattr(lib, f"array_max_idx_{sfx}")(V.data_ptr(), size, stride, ctypes.byref(idx))
        assert idx.value == int(np.argmax(v[::stride][:size]))


def test_every_kernel_sys_symbol_was_called(lib):
    """Bookkeeping for the three tests above: their name patterns cover all 126 reference symbols."""
    _, ref = lib
    fams = ["array_array_{op}_{s}", "array_array_{op}_assign_{s}", "array_scalar_{op}_{s}", "array_scalar_{op}_assign_{s}",
            "array_scalar_pointer_{op}_{s}", "array_scalar_pointer_{op}_assign_{s}"]
    called = {f.format(op=op, s=s) for f in fams for op in ("add", "sub", "mul", "div") for s in ("float", "double")}
    for u in ("sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "abs", "sqrt", "exp", "log"):
        called |= {f"array_{u}_{s}" for s in ("float", "double")} | {f"array_{u}_assign_{s}" for s in ("float", "double")}
    for s in ("float", "double"):
        called |= {f"array_clip_{s}", f"array_clip_assign_{s}", f"array_clip_backward_{s}", f"array_clip_backward_assign_{s}",
                   f"array_pow_{s}", f"array_pow_assign_{s}", f"relu_{s}", f"relu_backward_mask_{s}", f"conv_bias_add_{s}",
                   f"conv2d_bias_bkwd_{s}", f"memory_access_{s}", f"memory_set_{s}", f"array_max_idx_{s}"}
    assert called == set(ref), sorted(set(ref) ^ called)


# ---------------------------------------------------------------------------------------------- cuDNN-frontend wrapper shim
class Shape(ctypes.Structure):
    _fields_ = [("num_dims", ctypes.c_size_t), ("dims", ctypes.c_int64 * 8), ("strides", ctypes.c_int64 * 8)]


class Info(ctypes.Structure):
    _fields_ = [("padding", ctypes.c_int64 * 2), ("stride", ctypes.c_int64 * 2), ("dilation", ctypes.c_int64 * 2), ("num_dims", ctypes.c_int64)]


def shp(dims, strides=None):
    s = Shape()
    s.num_dims = len(dims)
    st = 1
    for i in range(len(dims) - 1, -1, -1):
        s.dims[i] = dims[i]
        s.strides[i] = st if strides is None else strides[i]
        st *= dims[i]
    return s


@pytest.mark.parametrize("dt,case", [(1, (2, 32, 12, 12, 64, 3, 1, 1)), (1, (2, 8, 9, 11, 6, 3, 1, 2)), (2, (2, 4, 8, 8, 6, 3, 0, 1))])
def test_fe_wrapper_conv_forward_backward_data_backward_filter(lib, dt, case):
    """create -> check -> workspace -> execute for all three conv descriptors (cudnn_frontend_wrapper.h:100-186; the reference's own
    C++ test only checks the status codes, tests/conv.cpp:72-133), against the oracle."""
    lib, _ = lib
    T = np.float32 if dt == 1 else np.float64
    n, c, h, w, k, r, pad, stride = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((n, c, h, w)).astype(T)
    wt = (rng.standard_normal((k, c, r, r)) * 0.2).astype(T)
    y_ref = zo.conv2d_fwd(x.astype(np.float64), wt.astype(np.float64), pad, stride, 1)
    dy = rng.standard_normal(y_ref.shape).astype(T)
    tol = 1e-3 if dt == 1 else 1e-10
    info = Info()
    info.padding[:] = [pad, pad]; info.stride[:] = [stride, stride]; info.dilation[:] = [1, 1]; info.num_dims = 2
    xs, ws, ys = shp(x.shape), shp(wt.shape), shp(y_ref.shape)
    vp = ctypes.c_void_p

    class B3(ctypes.Structure):
        _fields_ = [("a", vp), ("b", vp), ("c", vp)]

    def run(create, check, wsize, execute, destroy, shapes, bufs):
        desc = vp()
        fn = getattr(lib, create)
        fn.argtypes = [vp, ctypes.c_int] + [vp] * 4
        assert fn(ctypes.byref(desc), dt, *[ctypes.byref(s) for s in shapes], ctypes.byref(info)) == 0
        getattr(lib, check).argtypes = [vp, vp]
        assert getattr(lib, check)(desc, None) == 0
        size = ctypes.c_int64(-1)
        getattr(lib, wsize).argtypes = [vp, vp]
        assert getattr(lib, wsize)(desc, ctypes.byref(size)) == 0 and size.value == 0
        b = B3(*[t.data_ptr() for t in bufs])
        getattr(lib, execute).argtypes = [vp] * 4
        assert getattr(lib, execute)(desc, ctypes.byref(b), None, None) == 0
        getattr(lib, destroy).argtypes = [vp]
        getattr(lib, destroy)(desc)

    def rel(a, b):
        return float(np.linalg.norm((a - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-300))

    X, W, DY = dev(x), dev(wt), dev(dy)
    Y = dev(np.zeros(y_ref.shape, T))
    run("create_conv_descriptor", "check_conv_graph", "get_conv_workspace_size", "execute_conv_forward", "destroy_conv_descriptor",
        (xs, ws, ys), (X, W, Y))                                                        # ConvBufers {X, filter, Y}
    assert rel(host(Y), y_ref) < tol
    DX = dev(np.zeros_like(x))
    run("create_conv_backward_data_descriptor", "check_conv_backward_data_graph", "get_conv_backward_data_workspace_size",
        "execute_conv_backward_data", "destroy_conv_backward_data_descriptor", (ys, ws, xs), (DY, W, DX))   # {DY, filter, DX}
    assert rel(host(DX), zo.conv2d_bkwd_data(dy.astype(np.float64), wt.astype(np.float64), x.shape, pad, stride, 1)) < tol
    DW = dev(np.zeros_like(wt))
    run("create_conv_backward_filter_descriptor", "check_conv_backward_filter_graph", "get_conv_backward_filter_workspace_size",
        "execute_conv_backward_filter", "destroy_conv_backward_filter_descriptor", (xs, ys, ws), (X, DY, DW))   # {X, DY, DW}
    assert rel(host(DW), zo.conv2d_bkwd_filter(dy.astype(np.float64), x.astype(np.float64), wt.shape, pad, stride, 1)) < tol


@pytest.mark.parametrize("dt", [1, 2])
@pytest.mark.parametrize("nhwc", [False, True])
def test_fe_wrapper_batch_norm_forward_backward(lib, dt, nhwc):
    """The nine BatchNorm entry points (cudnn_frontend_wrapper.h:33-98; Rust caller zenu-cuda/src/cudnn/graph_batchnorm.rs:17-190):
    forward-training with cuDNN-frontend's conventions -- epsilon from the descriptor, next_running = (1 - momentum) * prev +
    momentum * batch, unbiased running variance, saved mean / inverse std -- and backward, against the oracle's formulas."""
    lib, _ = lib
    T = np.float32 if dt == 1 else np.float64
    n, c, h, w = 4, 8, 5, 6
    rng = np.random.default_rng(5 + dt)
    x = (rng.standard_normal((n, c, h, w)) * 1.5 + 0.5).astype(T)
    dy = rng.standard_normal((n, c, h, w)).astype(T)
    scale = rng.uniform(0.5, 1.5, c).astype(T)
    bias = rng.standard_normal(c).astype(T)
    rm0, rv0 = rng.standard_normal(c).astype(T), rng.uniform(0.5, 2.0, c).astype(T)
    eps, mom = 1e-3, 0.25
    cnt = n * h * w
    x64 = x.astype(np.float64)
    mean = x64.mean(axis=(0, 2, 3))
    var = x64.var(axis=(0, 2, 3))
    inv = 1.0 / np.sqrt(var + eps)
    y_ref = (x64 - mean[None, :, None, None]) * inv[None, :, None, None] * scale[None, :, None, None] + bias[None, :, None, None]
    rm_ref = (1 - mom) * rm0 + mom * mean
    rv_ref = (1 - mom) * rv0 + mom * var * cnt / (cnt - 1)
    dx_ref, ds_ref, db_ref = zo.bn2d_bwd(x64, dy.astype(np.float64), scale.astype(np.float64), mean, inv)
    if nhwc:
        to_dev = lambda a: dev(np.transpose(a, (0, 2, 3, 1)))  # noqa: E731
        back = lambda t: np.transpose(host(t), (0, 3, 1, 2))  # noqa: E731
        s = shp((n, c, h, w), strides=(h * w * c, 1, w * c, c))
    else:
        to_dev, back, s = dev, host, shp((n, c, h, w))
    vp = ctypes.c_void_p

    class FwdBufs(ctypes.Structure):
        _fields_ = [(k, vp) for k in ("X", "mean", "inv_variance", "scale", "bias", "peer_stats_0", "peer_stats_1", "prev_running_mean",
                                      "prev_running_var", "next_running_mean", "next_running_var", "Y")]

    class BwdBufs(ctypes.Structure):
        _fields_ = [(k, vp) for k in ("X", "DY", "scale", "mean", "inv_variance", "dscale", "dbias", "DX", "peer_stats_0", "peer_stats_1")]

    X, DY = to_dev(x), to_dev(dy)
    Y, DX = torch.zeros_like(X), torch.zeros_like(X)
    S, Bi, RM0, RV0 = dev(scale), dev(bias), dev(rm0), dev(rv0)
    SM, SI, RM1, RV1, DS, DB = (torch.zeros(c, dtype=X.dtype, device="cuda") for _ in range(6))
    desc = vp()
    lib.create_batch_norm_descriptor.argtypes = [vp, ctypes.c_int, vp, ctypes.c_float, ctypes.c_float, ctypes.c_bool]
    assert lib.create_batch_norm_descriptor(ctypes.byref(desc), dt, ctypes.byref(s), eps, mom, True) == 0
    lib.check_graph.argtypes = [vp, vp]
    assert lib.check_graph(desc, None) == 0
    size = ctypes.c_int64(-1)
    lib.get_workspace_size.argtypes = [vp, vp]
    assert lib.get_workspace_size(desc, ctypes.byref(size)) == 0 and size.value == 0
    lib.batch_norm_desc_debug.argtypes = [vp]
    lib.batch_norm_desc_debug(desc)
    fb = FwdBufs(X.data_ptr(), SM.data_ptr(), SI.data_ptr(), S.data_ptr(), Bi.data_ptr(), None, None, RM0.data_ptr(), RV0.data_ptr(),
                 RM1.data_ptr(), RV1.data_ptr(), Y.data_ptr())
    lib.execute_batch_norm_forward_training.argtypes = [vp] * 4
    assert lib.execute_batch_norm_forward_training(desc, ctypes.byref(fb), None, None) == 0
    lib.destroy_batch_norm_descriptor.argtypes = [vp]
    lib.destroy_batch_norm_descriptor(desc)
    tol = 2e-5 if dt == 1 else 1e-7      # (the descriptor carries epsilon / momentum as C floats)
    np.testing.assert_allclose(back(Y), y_ref, rtol=tol, atol=tol)
    np.testing.assert_allclose(host(SM), mean, rtol=tol, atol=tol)
    np.testing.assert_allclose(host(SI), inv, rtol=tol)
    np.testing.assert_allclose(host(RM1), rm_ref, rtol=tol, atol=tol)
    np.testing.assert_allclose(host(RV1), rv_ref, rtol=tol)
    np.testing.assert_array_equal(host(RM0), rm0)                 # prev_* are inputs
    bdesc = vp()
    lib.create_batch_norm_backward_data_descriptor.argtypes = [vp, ctypes.c_int, vp]
    assert lib.create_batch_norm_backward_data_descriptor(ctypes.byref(bdesc), dt, ctypes.byref(s)) == 0
    lib.check_backward_data_graph.argtypes = [vp, vp]
    assert lib.check_backward_data_graph(bdesc, None) == 0
    lib.get_backward_data_workspace_size.argtypes = [vp, vp]
    assert lib.get_backward_data_workspace_size(bdesc, ctypes.byref(size)) == 0 and size.value == 0
    bb = BwdBufs(X.data_ptr(), DY.data_ptr(), S.data_ptr(), SM.data_ptr(), SI.data_ptr(), DS.data_ptr(), DB.data_ptr(), DX.data_ptr(), None, None)
    lib.execute_batch_norm_backward_data.argtypes = [vp] * 4
    assert lib.execute_batch_norm_backward_data(bdesc, ctypes.byref(bb), None, None) == 0
    lib.destroy_batch_norm_backward_data_descriptor.argtypes = [vp]
    lib.destroy_batch_norm_backward_data_descriptor(bdesc)
    btol = 2e-4 if dt == 1 else 1e-6
    np.testing.assert_allclose(back(DX), dx_ref, rtol=btol, atol=btol)
    np.testing.assert_allclose(host(DS), ds_ref, rtol=btol, atol=btol)
    np.testing.assert_allclose(host(DB), db_ref, rtol=btol, atol=btol)
    # strides this library does not serve are refused at check time, like an unsupported cuDNN graph
    bad = shp((n, c, h, w), strides=(2 * c * h * w, h * w, w, 1))
    d2 = vp()
    assert lib.create_batch_norm_descriptor(ctypes.byref(d2), dt, ctypes.byref(bad), eps, mom, True) == 0
    assert lib.check_graph(d2, None) == 3   # NOT_SUPPORTED
    lib.destroy_batch_norm_descriptor(d2)
