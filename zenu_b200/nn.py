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

    def save_state(self, path):
        """Parameters + optimizer state (step count, Adam m / v): the file a data-parallel job resumes from."""
        check(self.lib.zb_model_save_state(self._h, str(path).encode()))

    def load_state(self, path):
        check(self.lib.zb_model_load_state(self._h, str(path).encode()))

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

    def train_step_async(self, x, targets, loss_out):
        """`train_step` without the host waiting for it: the step is enqueued and its loss lands in a pinned two-slot ring behind it
        (`loss_wait`).  `loss_out`: the persistent device scalar the step writes its loss to."""
        n, c, h, w = x.shape
        check(self.lib.zb_model_train_step_async(self._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(targets.data_ptr()), n, c, h, w,
                                                 ctypes.c_void_p(loss_out.data_ptr())))

    def loss_wait(self, age=0):
        """Loss of the step enqueued `age` `train_step_async` calls ago (0 = the latest, 1 = the one before); blocks until that step is done."""
        host = ctypes.c_double(0.0)
        check(self.lib.zb_model_loss_wait(self._h, int(age), ctypes.byref(host)))
        return host.value

    def set_wgrad_overlap(self, enable=True):
        """conv wgrad on the ctx's side stream, overlapping the following layers' BatchNorm backward (bit-identical results)."""
        check(self.lib.zb_model_set_wgrad_overlap(self._h, int(bool(enable))))

    def set_graph(self, enable=True):
        """Replay `train_step` from a CUDA graph: each distinct (buffers, batch shape, train mode, math mode, DP world) signature is
        captured after two eager steps of its own -- SGD / Adam / AdamW, bucket allreduces included -- and dropped when the buffers
        it bakes in go away.  Stays eager on the legacy default stream (not capturable) and while per-node profiling is on."""
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


class InputStage:
    """Library-owned input staging (zb_input_stage_*): pinned multi-buffered uint8 batches + int32 labels, asynchronous
    host->device copy of the BYTES on a copy stream, on-device expansion into the model's NCHW float batch and one-hot targets.

        stage = InputStage(ctx, n, c, h, w, classes, mean, std)
        img, lab = stage.host_buffers(slot)        # numpy views of the pinned buffers: the decoder writes into them
        stage.submit(slot)                         # H2D on the copy stream (overlaps the step of the previous batch)
        x, t = stage.wait(slot)                    # compute stream waits, expands; torch views of the slot's device tensors

    Replaces zenu/src/dataset.rs:74-100 + the synchronous f32 copy of Matrix::to::<Nvidia>() (zenu-matrix/src/matrix.rs:139-160,486)."""

    def __init__(self, ctx, n, c, h, w, classes, mean=None, std=None, slots=2, src_layout=_lib.ZB_NHWC, dtype=torch.float32):
        import numpy as np
        self._np = np
        self.ctx, self.lib = ctx, ctx.lib
        self.n, self.c, self.h, self.w, self.classes, self.slots = int(n), int(c), int(h), int(w), int(classes), int(slots)
        self.src_layout, self.dtype = src_layout, dtype
        zdt = ZB_F32 if dtype == torch.float32 else ZB_F64
        cm = (ctypes.c_double * c)(*mean) if mean is not None else None
        cs = (ctypes.c_double * c)(*std) if std is not None else None
        self._h = ctypes.c_void_p()
        check(self.lib.zb_input_stage_create(ctx.handle, zdt, src_layout, self.n, self.c, self.h, self.w, self.classes, cm, cs,
                                             self.slots, ctypes.byref(self._h)))
        ctx._children.add(self)
        self.h2d_bytes = int(self.lib.zb_input_stage_h2d_bytes(self._h))

    def close(self):
        if self._h and self.ctx.handle:
            self.lib.zb_input_stage_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def host_buffers(self, slot):
        """(images uint8 [n,h,w,c] or [n,c,h,w] per src_layout, labels int32 [n]): numpy views of slot's PINNED host memory.
        Blocks until an earlier copy out of this slot has finished (zb_input_stage_host_sync)."""
        np = self._np
        check(self.lib.zb_input_stage_host_sync(self._h, int(slot)))
        img, lab = ctypes.c_void_p(), ctypes.c_void_p()
        check(self.lib.zb_input_stage_host_buffers(self._h, int(slot), ctypes.byref(img), ctypes.byref(lab)))
        shape = (self.n, self.h, self.w, self.c) if self.src_layout == _lib.ZB_NHWC else (self.n, self.c, self.h, self.w)
        nbytes = self.n * self.c * self.h * self.w
        a = np.frombuffer((ctypes.c_uint8 * nbytes).from_address(img.value), dtype=np.uint8).reshape(shape)
        b = np.frombuffer((ctypes.c_int32 * self.n).from_address(lab.value), dtype=np.int32)
        return a, b

    def submit(self, slot):
        check(self.lib.zb_input_stage_submit(self._h, int(slot)))

    def wait(self, slot):
        x, t = ctypes.c_void_p(), ctypes.c_void_p()
        check(self.lib.zb_input_stage_wait(self._h, int(slot), ctypes.byref(x), ctypes.byref(t)))
        typestr = "<f4" if self.dtype == torch.float32 else "<f8"
        dev = f"cuda:{self.ctx.device}"
        xs = torch.as_tensor(_DevView(x.value, (self.n, self.c, self.h, self.w), typestr), device=dev)
        ts = torch.as_tensor(_DevView(t.value, (self.n, self.classes), typestr), device=dev)
        return xs, ts
