"""Python face of the C++ host model (zenu_b200/csrc/host, C ABI `zb_model_*`).

`Model` mirrors how the reference is driven (zenu/examples/mnist.rs:126-142):
    pred = model.call(x); loss = cross_entropy(pred, t); loss.backward(); optimizer.update(&model); loss.clear_grad()
collapsed into `forward_backward()` + `update()` (= `train_step()`), all executed by the native library.
Parameters are exposed under the reference's names as zero-copy torch views of the flat device buffers.
"""
import ctypes

import torch

from . import _lib
from ._lib import ZB_F32, ZB_F64, check

OPT = {"sgd": 0, "adam": 1, "adamw": 2}
KIND = {0: "weight", 1: "bias", 2: "buffer"}


class _DevView:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned device memory."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class Model:
    def __init__(self, ctx, arch, num_classes, dtype=torch.float32, fused=True, seed=42, bucket_mb=25):
        self.ctx = ctx
        self.lib = ctx.lib
        self.arch = arch
        self.num_classes = int(num_classes)
        self.dtype = dtype
        self._zdt = ZB_F32 if dtype == torch.float32 else ZB_F64
        self._h = ctypes.c_void_p()
        check(self.lib.zb_model_create(ctx.handle, arch.encode(), self._zdt, self.num_classes, int(bool(fused)), int(seed),
                                       int(bucket_mb) << 20, ctypes.byref(self._h)))
        self._params = None
        ctx._children.add(self)

    def close(self):
        if self._h and self.ctx.handle:
            self.lib.zb_model_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Parameters::parameters() ---------------------------------------------------------------------------------
    def named_parameters(self):
        """name -> dict(data=tensor view, grad=tensor view or None, kind='weight'|'bias'|'buffer').
        Conv filters are KRSC ([K,R,S,C]); use `filter_to_kcrs` for the reference layout."""
        if self._params is None:
            out = {}
            n = self.lib.zb_model_param_count(self._h)
            typestr = "<f4" if self.dtype == torch.float32 else "<f8"
            dev = f"cuda:{self.ctx.device}"
            for i in range(n):
                name = ctypes.create_string_buffer(256)
                shape = (ctypes.c_int64 * 4)()
                ndim, kind = ctypes.c_int(), ctypes.c_int()
                data, grad = ctypes.c_void_p(), ctypes.c_void_p()
                check(self.lib.zb_model_param_info(self._h, i, name, 256, shape, ctypes.byref(ndim), ctypes.byref(kind),
                                                   ctypes.byref(data), ctypes.byref(grad)))
                shp = [shape[j] for j in range(ndim.value)]
                d = torch.as_tensor(_DevView(data.value, shp, typestr), device=dev)
                g = torch.as_tensor(_DevView(grad.value, shp, typestr), device=dev) if grad.value else None
                out[name.value.decode()] = {"data": d, "grad": g, "kind": KIND[kind.value]}
            self._params = out
        return self._params

    @staticmethod
    def filter_to_kcrs(t):
        return t.permute(0, 3, 1, 2).contiguous()

    @staticmethod
    def filter_from_kcrs(t):
        return t.permute(0, 2, 3, 1).contiguous()

    def train(self, flag=True):
        check(self.lib.zb_model_set_train(self._h, int(bool(flag))))

    # ---- zenu::save_model / load_model (reference file format, zenu/src/lib.rs:26-67) ------------------------------------
    def save(self, path):
        check(self.lib.zb_model_save(self._h, str(path).encode()))

    def load(self, path):
        check(self.lib.zb_model_load(self._h, str(path).encode()))

    def set_optimizer(self, kind="sgd", lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
        check(self.lib.zb_model_set_optimizer(self._h, OPT[kind], float(lr), float(beta1), float(beta2), float(eps),
                                              float(weight_decay)))

    # ---- Module::call ------------------------------------------------------------------------------------------------
    def forward(self, x):
        n, c, h, w = x.shape
        out = torch.empty((n, self.num_classes), dtype=self.dtype, device=x.device)
        check(self.lib.zb_model_forward(self._h, ctypes.c_void_p(x.data_ptr()), n, c, h, w, ctypes.c_void_p(out.data_ptr())))
        return out

    def forward_backward(self, x, targets, loss_out=None):
        n, c, h, w = x.shape
        loss = loss_out if loss_out is not None else torch.empty((1,), dtype=self.dtype, device=x.device)
        check(self.lib.zb_model_forward_backward(self._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(targets.data_ptr()),
                                                 n, c, h, w, ctypes.c_void_p(loss.data_ptr())))
        return loss

    def update(self):
        check(self.lib.zb_model_update(self._h))

    def train_step(self, x, targets, loss_out=None, read_loss=False):
        """One optimisation step.  With read_loss the scalar loss is copied to the host (synchronises)."""
        n, c, h, w = x.shape
        if loss_out is not None:
            loss = loss_out
        elif getattr(self, "_graph", False):   # graph replay is keyed by the buffer addresses: one persistent loss scalar
            if getattr(self, "_loss_buf", None) is None:
                self._loss_buf = torch.empty((1,), dtype=self.dtype, device=x.device)
            loss = self._loss_buf
        else:
            loss = torch.empty((1,), dtype=self.dtype, device=x.device)
        host = ctypes.c_double(0.0)
        check(self.lib.zb_model_train_step(self._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(targets.data_ptr()), n, c, h, w,
                                           ctypes.c_void_p(loss.data_ptr()), ctypes.byref(host) if read_loss else None))
        return host.value if read_loss else loss

    def set_graph(self, enable=True):
        """Replay `train_step` from a CUDA graph (captured after two eager steps; single GPU + SGD, otherwise stays eager)."""
        check(self.lib.zb_model_set_graph(self._h, int(bool(enable))))
        self._graph = bool(enable)

    def graph_count(self):
        """Number of step graphs captured so far (0 = every step ran eagerly)."""
        return int(self.lib.zb_model_graph_count(self._h))

    def profile(self, enable=True):
        """Per-node CUDA-event timing of the tape (forward and backward nodes), see `profile_table`."""
        check(self.lib.zb_model_profile_enable(self._h, int(bool(enable))))

    def profile_table(self):
        """[(key, count, total_ms, algorithmic_flops, algorithmic_bytes)] accumulated since profile(True)."""
        need = int(self.lib.zb_model_profile_dump(self._h, None, 0))
        buf = ctypes.create_string_buffer(need + 16)
        self.lib.zb_model_profile_dump(self._h, buf, need + 16)
        rows = []
        for ln in buf.value.decode().splitlines():
            k, n, ms, fl, by = ln.split("\t")
            rows.append((k, int(n), float(ms), float(fl), float(by)))
        return rows

    def bytes_reserved(self):
        return int(self.lib.zb_model_bytes_reserved(self._h))
