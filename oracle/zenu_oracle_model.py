"""CPU reference train step for the configs' models, composed from the oracle ops (oracle/zenu_oracle.py).

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline / --impl reference).

It replays what the reference's tape would execute on its CPU device for
  small_cnn  (zenu/examples/cifar10.rs:29-69),  resnet18 / resnet50 (torchvision v1.5 topology built from
  zenu-layer's Conv2d / BatchNorm2d / Linear as in the ResBlock sketch zenu/examples/resnet.rs:18-28):
per layer  conv (im2col+gemm) [+bias] -> batch_norm_2d -> relu [-> add] ... -> cross_entropy, then the reverse
sweep in the order Variable::backward visits the nodes (zenu-autograd/src/lib.rs:220-237) and
Optimizer::update (zenu-optimizer/src/{sgd,adam,adamw}.rs).  Tensors are NCHW, filters KCRS, parameter names
are the reference's ("conv1.conv2d.filter", "bn1.batch_norm_2d.scale", ...).
"""
import numpy as np

from . import zenu_oracle as zo


def _resnet_plan(depth):
    bottleneck = depth == 50
    counts = [3, 4, 6, 3] if bottleneck else [2, 2, 2, 2]
    blocks = []
    cin = 64
    for stage, cnt in enumerate(counts):
        width = 64 << stage
        for b in range(cnt):
            stride = 2 if (b == 0 and stage > 0) else 1
            cout = width * 4 if bottleneck else width
            name = f"layer{stage + 1}.{b}"
            if bottleneck:
                convs = [(cin, width, 1, 1, 0), (width, width, 3, stride, 1), (width, cout, 1, 1, 0)]
            else:
                convs = [(cin, width, 3, stride, 1), (width, cout, 3, 1, 1)]
            down = (cin, cout, 1, stride, 0) if (stride != 1 or cin != cout) else None
            blocks.append((name, convs, down))
            cin = cout
    return blocks, cin


def param_shapes(arch, num_classes):
    """Ordered {name: (shape, kind)} in forward order; filters KCRS."""
    out = {}

    def conv(p, ci, co, k, bias=False):
        out[p + ".conv2d.filter"] = ((co, ci, k, k), "weight")
        if bias:
            out[p + ".conv2d.bias"] = ((co,), "bias")

    def bn(p, c):
        out[p + ".batch_norm_2d.scale"] = ((c,), "weight")
        out[p + ".batch_norm_2d.bias"] = ((c,), "bias")
        out[p + ".batch_norm_2d.mean"] = ((c,), "buffer")
        out[p + ".batch_norm_2d.variance"] = ((c,), "buffer")

    def lin(p, i, o):
        out[p + ".linear.weight"] = ((o, i), "weight")
        out[p + ".linear.bias"] = ((o,), "bias")

    if arch == "small_cnn":
        conv("conv1", 3, 32, 3, True); bn("batch_norm1", 32)
        conv("conv2", 32, 64, 3, True); bn("batch_norm2", 64)
        lin("linear1", 64 * 32 * 32, 512); lin("linear2", 512, num_classes)
    else:
        blocks, cfin = _resnet_plan(18 if arch == "resnet18" else 50)
        conv("conv1", 3, 64, 7); bn("bn1", 64)
        for name, convs, down in blocks:
            for i, (ci, co, k, _, _) in enumerate(convs):
                conv(f"{name}.conv{i + 1}", ci, co, k); bn(f"{name}.bn{i + 1}", co)
            if down:
                conv(f"{name}.downsample_conv", down[0], down[1], 1); bn(f"{name}.downsample_bn", down[1])
        lin("fc", cfin, num_classes)
    return out


def init_params(arch, num_classes, seed=42, dtype=np.float32):
    """He-normal filters, N(0,1)/sqrt(in) Linear (linear.rs:56-60), BN scale 1 / bias 0 / mean 0 / var 1."""
    rng = np.random.default_rng(seed)
    p = {}
    for name, (shape, kind) in param_shapes(arch, num_classes).items():
        if name.endswith("conv2d.filter"):
            fan = shape[1] * shape[2] * shape[3]
            p[name] = (rng.standard_normal(shape) * np.sqrt(2.0 / fan)).astype(dtype)
        elif name.endswith("linear.weight"):
            p[name] = (rng.standard_normal(shape) / np.sqrt(shape[1])).astype(dtype)
        elif name.endswith("scale") or name.endswith("variance"):
            p[name] = np.ones(shape, dtype)
        else:
            p[name] = np.zeros(shape, dtype)
    return p


class OracleModel:
    def __init__(self, arch, num_classes, params, momentum=0.9, operand_round=None):
        """operand_round: None = the reference's f32 arithmetic; "rna"/"rne"/"rz" = additionally round every conv / GEMM
        OPERAND to tf32 first (zo.tf32_round), modelling the device's tensor-core math mode for model-level tests."""
        self.arch, self.num_classes, self.p, self.momentum = arch, num_classes, params, momentum
        self.operand_round = operand_round
        self.rnd = (lambda a: zo.tf32_round(a, operand_round)) if operand_round else (lambda a: a)
        self.dtype = next(iter(params.values())).dtype
        self.kinds = {k: v[1] for k, v in param_shapes(arch, num_classes).items()}
        self.state = {}
        self.step = 0

    # ---- forward pieces; each returns (output, backward closure) ---------------------------------------------------
    def _conv(self, name, x, stride, pad, grads):
        w = self.p[name + ".conv2d.filter"]
        b = self.p.get(name + ".conv2d.bias")
        rnd = self.rnd
        y = zo.conv2d_fwd(rnd(x), rnd(w), pad, stride, 1)
        if b is not None:
            y = zo.conv2d_bias_add(y, b)

        def back(dy):
            if b is not None:
                grads[name + ".conv2d.bias"] = zo.conv2d_bias_bkwd(dy)
            grads[name + ".conv2d.filter"] = zo.conv2d_bkwd_filter(rnd(dy), rnd(x), w.shape, pad, stride, 1)
            return zo.conv2d_bkwd_data(rnd(dy), rnd(w), x.shape, pad, stride, 1)
        return y, back

    def _bn(self, name, x, grads):
        sc, bi = self.p[name + ".batch_norm_2d.scale"], self.p[name + ".batch_norm_2d.bias"]
        y, rm, rv, sm, si = zo.bn2d_fwd_train(x, sc, bi, self.p[name + ".batch_norm_2d.mean"],
                                               self.p[name + ".batch_norm_2d.variance"], self.momentum)
        self.p[name + ".batch_norm_2d.mean"], self.p[name + ".batch_norm_2d.variance"] = rm, rv

        def back(dy):
            dx, ds, db = zo.bn2d_bwd(x, dy, sc, sm, si)
            grads[name + ".batch_norm_2d.scale"], grads[name + ".batch_norm_2d.bias"] = ds, db
            return dx
        return y, back

    @staticmethod
    def _relu(x):
        y = zo.relu(x)
        return y, (lambda dy: zo.ewise("mul", dy, zo.relu_backward_mask(x)))

    def _linear(self, name, x, grads):
        w, b = self.p[name + ".linear.weight"], self.p[name + ".linear.bias"]
        rnd = self.rnd
        y = zo.linear_fwd(rnd(x), rnd(w), b)

        def back(dy):
            dx, dw, db = zo.linear_bwd(rnd(x), rnd(w), rnd(dy))
            if self.operand_round:
                db = zo.linear_bwd(x, w, dy)[2]   # the bias gradient is a plain f32 column sum on the device
            grads[name + ".linear.weight"], grads[name + ".linear.bias"] = dw, db
            return dx
        return y, back

    def forward_backward(self, x, t):
        """Returns (loss, grads)."""
        x = np.ascontiguousarray(x, self.dtype)
        t = np.ascontiguousarray(t, self.dtype)
        grads, tape = {}, []

        def seq(fn_out):
            y, back = fn_out
            tape.append(back)
            return y

        if self.arch == "small_cnn":
            h = seq(self._conv("conv1", x, 1, 1, grads)); h = seq(self._bn("batch_norm1", h, grads)); h = seq(self._relu(h))
            h = seq(self._conv("conv2", h, 1, 1, grads)); h = seq(self._bn("batch_norm2", h, grads)); h = seq(self._relu(h))
            shp = h.shape
            h = h.reshape(shp[0], -1); tape.append(lambda d: d.reshape(shp))
            h = seq(self._linear("linear1", h, grads)); h = seq(self._relu(h))
            logits = seq(self._linear("linear2", h, grads))
        else:
            blocks, _ = _resnet_plan(18 if self.arch == "resnet18" else 50)
            h = seq(self._conv("conv1", x, 2, 3, grads)); h = seq(self._bn("bn1", h, grads)); h = seq(self._relu(h))
            xin = h
            h = zo.maxpool2d_fwd(xin, 3, 2, 1)
            tape.append(lambda d, xin=xin: zo.maxpool2d_bwd(xin, d, 3, 2, 1))
            for name, convs, down in blocks:
                h = seq(self._block(name, convs, down, h, grads))
            hw = h.shape[2:]
            h = zo.gap_fwd(h); tape.append(lambda d, hw=hw: zo.gap_bwd(d, hw))
            logits = seq(self._linear("fc", h, grads))
        loss, dz = zo.softmax_xent(logits, t)
        d = dz
        for back in reversed(tape):
            d = back(d)
        return loss, grads

    def _block(self, name, convs, down, x, grads):
        sub = []
        h = x
        for i, (_, _, _, stride, pad) in enumerate(convs):
            y, b1 = self._conv(f"{name}.conv{i + 1}", h, stride, pad, grads); sub.append(b1)
            y, b2 = self._bn(f"{name}.bn{i + 1}", y, grads); sub.append(b2)
            if i < len(convs) - 1:
                y, b3 = self._relu(y); sub.append(b3)
            h = y
        dsub = []
        sc = x
        if down:
            sc, d1 = self._conv(f"{name}.downsample_conv", x, down[3], 0, grads); dsub.append(d1)
            sc, d2 = self._bn(f"{name}.downsample_bn", sc, grads); dsub.append(d2)
        z = zo.ewise("add", h, sc)
        out, rb = self._relu(z)

        def back(dy):
            dz = rb(dy)
            dm = dz
            for b in reversed(sub):
                dm = b(dm)
            ds = dz
            for b in reversed(dsub):
                ds = b(ds)
            return zo.ewise("add", dm, ds)  # grad fan-in: grad + old (lib.rs:480-481)
        return out, back

    # ---- Optimizer::update ----------------------------------------------------------------------------------------------
    def update(self, grads, kind="sgd", lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
        self.step += 1
        for name, g in grads.items():
            p = self.p[name]
            if kind == "sgd":
                zo.sgd_step(p, g.reshape(p.shape), lr)
            else:
                m = self.state.setdefault(name + ".m", np.zeros_like(p))
                v = self.state.setdefault(name + ".v", np.zeros_like(p))
                zo.adam_step(p, g.reshape(p.shape), m, v, lr, beta1, beta2, eps, self.step, weight_decay,
                             kind == "adamw" and self.kinds[name] == "weight")

    def train_step(self, x, t, **opt):
        loss, grads = self.forward_backward(x, t)
        self.update(grads, **opt)
        return loss
