"""ctypes loader for libzenu_b200.so — the C-ABI boundary (include/zenu_b200.h).

The product path fails loudly when the CUDA library is missing: there is no CPU or PyTorch fallback
behind any op in this package.  Prototypes are derived from the public header so the binding cannot
drift from the ABI.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# ZENU_B200_LIB: another build of the same library (A/B of build-time knobs); it must still be the in-tree CUDA library
LIB_PATH = os.environ.get("ZENU_B200_LIB") or os.path.join(_HERE, "lib", "libzenu_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

ZB_OK = 0
ZB_F32, ZB_F64 = 0, 1
ZB_NCHW, ZB_NHWC, ZB_NCHW_X = 0, 1, 2
ZB_MATH_DEFAULT, ZB_MATH_TF32, ZB_MATH_TF32X3, ZB_MATH_FP32 = 0, 1, 2, 3
ZB_OP_ADD, ZB_OP_SUB, ZB_OP_MUL, ZB_OP_DIV = 0, 1, 2, 3
CUDA_STREAM_LEGACY = 0x1


class ZenuB200Error(RuntimeError):
    pass


class ConvDesc(ctypes.Structure):
    """zb_conv2d_desc"""
    _fields_ = [(n, ctypes.c_int64) for n in
                ("n", "c", "h", "w", "k", "kh", "kw", "pad_h", "pad_w", "stride_h", "stride_w", "dil_h", "dil_w")]


_CTYPE = {
    "int": ctypes.c_int, "int64_t": ctypes.c_int64, "double": ctypes.c_double, "float": ctypes.c_float,
    "unsigned long long": ctypes.c_ulonglong, "size_t": ctypes.c_size_t, "uint64_t": ctypes.c_uint64,
}


def _map_type(t):
    t = t.replace("const", "").strip()
    t = re.sub(r"\s+", " ", t)
    if t == "void":
        return None
    if t == "char*":
        return ctypes.c_char_p
    if t.endswith("*"):
        return ctypes.c_void_p
    if t in _CTYPE:
        return _CTYPE[t]
    if t in ("CudnnFrontendError_t", "CudnnFrontendDataType_t", "zb_status"):
        return ctypes.c_int
    raise KeyError(t)


def parse_header(path):
    """Return {name: (restype, [argtypes])} for every function declared in a C header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b([A-Za-z_]\w*)\s*\(([^()]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef") or not ret:
            continue
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)([A-Za-z_]\w*)?$", a)
                ty = mm.group(1).strip() if mm.group(1).strip() else a
                # "T* name" / "T *name" / "T name"
                ty = a[: a.rfind(mm.group(2))].strip() if mm.group(2) and not a.endswith("*") and a.rfind(mm.group(2)) > 0 else a
                argtypes.append(_map_type(ty))
        out[name] = (_map_type(ret), argtypes)
    return out


_lib = None
_protos = None


def load():
    """Load the shared library (built by __graft_entry__.build() / make -C zenu_b200/csrc)."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZenuB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). zenu_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header(os.path.join(INCLUDE_DIR, "zenu_b200.h"))
    for name, (ret, args) in _protos.items():
        fn = getattr(lib, name)  # AttributeError here = header/library drift; fail loudly
        fn.restype = ret
        fn.argtypes = args
    _lib = lib
    return lib


def prototypes():
    load()
    return _protos


def check(rc):
    if rc != ZB_OK:
        msg = load().zb_last_error()
        raise ZenuB200Error(f"zenu_b200 status {rc}: {msg.decode() if msg else ''}")
