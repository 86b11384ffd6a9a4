"""Cross-checks the oracle's model composition (oracle/zenu_oracle_model.py) against torch-CPU float64 autograd:
the cases no reference test pins (multi-layer tapes, stride-2 / 1x1 / 7x7 convs, residual fan-in, max-pool ties)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import zenu_oracle_model as zm


def torch_forward(arch, P, x, t):
    def conv(n, h, s, p):
        return F.conv2d(h, P[n + ".conv2d.filter"], P.get(n + ".conv2d.bias"), stride=s, padding=p)

    def bn(n, h):
        return F.batch_norm(h, None, None, P[n + ".batch_norm_2d.scale"], P[n + ".batch_norm_2d.bias"], True, 0.1, 1e-10)

    if arch == "small_cnn":
        h = F.relu(bn("batch_norm1", conv("conv1", x, 1, 1)))
        h = F.relu(bn("batch_norm2", conv("conv2", h, 1, 1)))
        h = h.flatten(1)
        h = F.relu(F.linear(h, P["linear1.linear.weight"], P["linear1.linear.bias"]))
        z = F.linear(h, P["linear2.linear.weight"], P["linear2.linear.bias"])
    else:
        blocks, _ = zm._resnet_plan(18 if arch == "resnet18" else 50)
        h = F.relu(bn("bn1", conv("conv1", x, 2, 3)))
        h = F.max_pool2d(F.pad(h, (1, 1, 1, 1)), 3, 2)   # zero padding takes part in the max (reference CPU semantics)
        for name, convs, down in blocks:
            sc = h
            y = h
            for i, (_, _, _, s, p) in enumerate(convs):
                y = bn(f"{name}.bn{i + 1}", conv(f"{name}.conv{i + 1}", y, s, p))
                if i < len(convs) - 1:
                    y = F.relu(y)
            if down:
                sc = bn(f"{name}.downsample_bn", conv(f"{name}.downsample_conv", h, down[3], 0))
            h = F.relu(y + sc)
        h = h.mean((2, 3))
        z = F.linear(h, P["fc.linear.weight"], P["fc.linear.bias"])
    return -(t * F.log_softmax(z, 1)).sum() / x.shape[0]


@pytest.mark.parametrize("arch,batch,hw,classes", [("small_cnn", 3, 32, 10), ("resnet18", 2, 32, 7), ("resnet50", 2, 32, 5)])
def test_oracle_model_matches_torch_f64(arch, batch, hw, classes):
    params = zm.init_params(arch, classes, seed=3, dtype=np.float64)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((batch, 3, hw, hw))
    t = np.zeros((batch, classes))
    t[np.arange(batch), rng.integers(0, classes, batch)] = 1.0
    model = zm.OracleModel(arch, classes, {k: v.copy() for k, v in params.items()})
    loss, grads = model.forward_backward(x, t)
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=("mean" not in k and "variance" not in k))
         for k, v in params.items()}
    tl = torch_forward(arch, P, torch.tensor(x), torch.tensor(t))
    tl.backward()
    assert abs(loss - tl.item()) < 1e-9 * max(1.0, abs(tl.item()))
    for k, g in grads.items():
        ref = P[k].grad.numpy()
        err = np.abs(g.reshape(ref.shape) - ref).max() / max(np.abs(ref).max(), 1e-3)  # conv bias before BN has a ~0 gradient
        assert err < 1e-6, (k, err)
    # running stats follow the reference rule: new = batch*(1-m) + old*m with m = 0.9
    first_bn = "bn1" if arch != "small_cnn" else "batch_norm1"
    assert not np.allclose(model.p[first_bn + ".batch_norm_2d.mean"], 0.0)


def test_oracle_sgd_training_reduces_loss():
    params = zm.init_params("small_cnn", 10, seed=1)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((8, 3, 32, 32)).astype(np.float32)
    t = np.zeros((8, 10), np.float32)
    t[np.arange(8), rng.integers(0, 10, 8)] = 1.0
    model = zm.OracleModel("small_cnn", 10, params)
    losses = [model.train_step(x, t, kind="sgd", lr=0.01) for _ in range(4)]
    assert losses[-1] < losses[0]
