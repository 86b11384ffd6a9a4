"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and exports every
symbol the three public headers declare (include/*.h)."""
import ctypes
import os
import re
import subprocess

import pytest

from zenu_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def _declared(header):
    pre = subprocess.run(["/usr/bin/gcc", "-E", "-P", "-x", "c", os.path.join(ROOT, "include", header)],
                         check=True, capture_output=True, text=True).stdout
    pre = re.sub(r"typedef[^;{]*\{[^}]*\}[^;]*;", "", pre, flags=re.S)
    names = re.findall(r"\b([A-Za-z_]\w*)\s*\([^()]*\)\s*;", pre)
    return sorted(set(n for n in names if not n.startswith("__")))


@pytest.mark.parametrize("header,minimum", [("zenu_b200.h", 40), ("zenu_kernel_compat.h", 120),
                                            ("zenu_cudnn_frontend_compat.h", 26)])
def test_exports_every_declared_symbol(lib, header, minimum):
    names = _declared(header)
    assert len(names) >= minimum, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"{header}: library does not export {missing}"


def _golden_tools():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("surface,header,count", [("kernel_sys", "zenu_kernel_compat.h", 126),
                                                  ("cudnn_frontend_wrapper", "zenu_cudnn_frontend_compat.h", 21)])
def test_reference_ffi_surface_is_exported_with_the_reference_prototypes(lib, surface, header, count):
    """The .so against the REFERENCE's headers, not this repo's own: tests/golden/ref_symbols.json holds every prototype of
    zenu-cuda-kernel-sys/kernel/*.h and cudnn_frontend_wrapper.h (extracted from /root/reference by tests/golden/make_golden.py).
    Every one of them must be exported, and the compat header must declare it with the same return type and parameter list
    (`cudnnHandle_t*` is carried as `void*`: there is no cuDNN behind this library), so the bindgen'd Rust crates
    (zenu-cuda-kernel-sys/build.rs:37-53, zenu-cudnn-frontend-wrapper-sys/build.rs) link and call unchanged."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "ref_symbols.json")) as f:
        ref = json.load(f)[surface]
    assert len(ref) == count
    missing = [n for n in ref if not hasattr(lib, n)]
    assert not missing, f"reference symbols the library does not export: {missing}"
    pre = subprocess.run(["/usr/bin/gcc", "-E", "-P", "-x", "c", os.path.join(ROOT, "include", header)],
                         check=True, capture_output=True, text=True).stdout   # the kernel-sys header is macro-generated
    ours = _golden_tools().c_prototypes(None, text=pre)
    norm = lambda t: t.replace("cudnnHandle_t*", "void*").replace("bool", "_Bool")  # noqa: E731  (stdbool.h: bool is _Bool after cpp)
    bad = []
    for name, proto in ref.items():
        mine = ours.get(name)
        if mine is None or norm(proto["ret"]) != mine["ret"] or [norm(a) for a in proto["args"]] != mine["args"]:
            bad.append((name, proto, mine))
    assert not bad, bad


def test_rust_sys_crate_in_sync():
    """rust/zenu-b200-sys/src/lib.rs (the bindings the Rust glue under rust/patches binds) is what tools/gen_rust_sys.py derives
    from include/zenu_b200.h today: every zb_* function, in header order, with the C types mapped 1:1."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_sys", os.path.join(ROOT, "tools", "gen_rust_sys.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    text = mod.generate()
    with open(mod.OUT) as f:
        assert f.read() == text, "run `python tools/gen_rust_sys.py` after changing include/zenu_b200.h"
    declared = set(_declared("zenu_b200.h"))
    bound = set(re.findall(r"pub fn (zb_\w+)\(", text))
    assert bound == declared, sorted(bound ^ declared)
    # every zb_* call made by the patched reference sources exists in the bindings
    used = set()
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rust", "patches")):
        for fn in files:
            used |= set(re.findall(r"sys::(zb_\w+)\(", open(os.path.join(dirpath, fn)).read()))
    assert used and used <= bound, sorted(used - bound)


def test_header_prototypes_parse():
    protos = _lib.parse_header(os.path.join(ROOT, "include", "zenu_b200.h"))
    assert protos["zb_conv2d_fprop"][1][4] is ctypes.c_void_p
    assert protos["zb_conv_out_size"][0] is ctypes.c_int64
    assert protos["zb_gemm"][1][8] is ctypes.c_double


def test_host_only_entry_points(lib):
    # pure host functions are safe without a GPU
    lib.zb_conv_out_size.restype = ctypes.c_int64
    lib.zb_conv_out_size.argtypes = [ctypes.c_int64] * 5
    assert lib.zb_conv_out_size(224, 7, 3, 2, 1) == 112   # ResNet conv1
    assert lib.zb_conv_out_size(56, 3, 1, 1, 1) == 56
    assert lib.zb_conv_out_size(32, 3, 1, 1, 2) == 30     # dilation 2
    lib.zb_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.zb_version()


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from zenu_b200 import ZenuB200Error, ops
    with pytest.raises(ZenuB200Error):
        ops.Context()


def test_product_never_imports_oracle():
    # the oracle is test infrastructure: nothing under zenu_b200/ may reference it
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zenu_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"\boracle\b|zenu_oracle|libzenu_oracle", src):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
