"""Model files in the reference's format (zenu::save_model / load_model, zenu/src/lib.rs:26-67): the bincode image of
HashMap<String, Variable>.  Thin ctypes binding of the host-only reader / writer in libzenu_b200.so (zb_ckpt_*); the
model-level save / load (device copies + KRSC <-> KCRS) is nn.Model.save / nn.Model.load (zb_model_save / zb_model_load)."""
import ctypes

import numpy as np

from . import _lib
from ._lib import ZB_F32, ZB_F64, ZenuB200Error


def _check(rc, lib):
    if rc != 0:
        raise ZenuB200Error(lib.zb_last_error().decode())


def write_state_dict(path, tensors):
    """tensors: {name: numpy f32 / f64 array} in the reference's layouts; all of one dtype (a reference model is Variable<T, D>)."""
    lib = _lib.load()
    names = list(tensors)
    arrs = [np.require(tensors[k], requirements=["C"]) for k in names]
    dts = {a.dtype for a in arrs}
    if len(dts) > 1 or (dts and next(iter(dts)) not in (np.dtype(np.float32), np.dtype(np.float64))):
        raise ZenuB200Error("write_state_dict: tensors must all be float32 or all float64")
    dtype = ZB_F64 if dts and next(iter(dts)) == np.dtype(np.float64) else ZB_F32
    n = len(names)
    c_names = (ctypes.c_char_p * n)(*[k.encode() for k in names])
    c_ndims = (ctypes.c_int * n)(*[a.ndim for a in arrs])
    shape_bufs = [(ctypes.c_int64 * max(a.ndim, 1))(*a.shape) for a in arrs]
    c_shapes = (ctypes.c_void_p * n)(*[ctypes.cast(b, ctypes.c_void_p) for b in shape_bufs])
    c_data = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrs])
    _check(lib.zb_ckpt_write(str(path).encode(), dtype, n, c_names, c_ndims, c_shapes, c_data), lib)


def read_state_dict(path):
    """-> {name: numpy array} (dense row-major copies; strides / ptr_offset of the file already resolved)."""
    lib = _lib.load()
    h = ctypes.c_void_p()
    _check(lib.zb_ckpt_open(str(path).encode(), ctypes.byref(h)), lib)
    out = {}
    try:
        for i in range(lib.zb_ckpt_count(h)):
            name = ctypes.create_string_buffer(512)
            shape = (ctypes.c_int64 * 8)()
            ndim, dtype = ctypes.c_int(), ctypes.c_int()
            data, numel = ctypes.c_void_p(), ctypes.c_int64()
            _check(lib.zb_ckpt_entry(h, i, name, 512, shape, ctypes.byref(ndim), ctypes.byref(dtype), ctypes.byref(data),
                                     ctypes.byref(numel)), lib)
            np_t = np.float64 if dtype.value == ZB_F64 else np.float32
            if numel.value:
                buf = (ctypes.c_char * (numel.value * np.dtype(np_t).itemsize)).from_address(data.value)
                arr = np.frombuffer(buf, dtype=np_t).copy()
            else:
                arr = np.zeros((0,), np_t)
            out[name.value.decode()] = arr.reshape([shape[k] for k in range(ndim.value)])
    finally:
        lib.zb_ckpt_close(h)
    return out
