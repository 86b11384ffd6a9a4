"""Reference model-file format (bincode image of HashMap<String, Variable>, zenu/src/lib.rs:26-67,
zenu-matrix/src/impl_serde.rs:11-40): host-only reader / writer checked against a hand-assembled byte image of the
reference's own serialisation fixture (impl_serde.rs test_matrix_serialization_format: shape [2,2], stride [2,1],
data [1,2,3,4], data_type "f32", ptr_offset 0), round trips, strided entries and error behaviour; model save / load on the GPU."""
import os
import struct

import numpy as np
import pytest

from zenu_b200 import ZenuB200Error, checkpoint


def _u64(v):
    return struct.pack("<Q", v)


def _s(txt):
    b = txt.encode()
    return _u64(len(b)) + b


def _entry(name, shape, stride, data, ty, ptr_offset=0):
    fmt = "<%d%s" % (len(data), "f" if ty == "f32" else "d")
    return (_s(name) + _u64(len(shape)) + b"".join(_u64(d) for d in shape) + _u64(len(stride)) + b"".join(_u64(d) for d in stride)
            + _u64(len(data)) + struct.pack(fmt, *data) + _s(ty) + _u64(ptr_offset))


def test_writer_matches_bincode_image_of_reference_fixture(tmp_path):
    path = tmp_path / "m.bin"
    checkpoint.write_state_dict(path, {"linear.weight": np.array([[1, 2], [3, 4]], np.float32)})
    expect = _u64(1) + _entry("linear.weight", [2, 2], [2, 1], [1.0, 2.0, 3.0, 4.0], "f32")
    assert path.read_bytes() == expect


def test_round_trip_f32_f64(tmp_path):
    rng = np.random.default_rng(0)
    for dt in (np.float32, np.float64):
        sd = {"conv1.conv2d.filter": rng.standard_normal((4, 3, 3, 3)).astype(dt), "conv1.conv2d.bias": rng.standard_normal((1, 4, 1, 1)).astype(dt),
              "bn1.batch_norm_2d.scale": rng.standard_normal((4,)).astype(dt), "empty": np.zeros((0,), dt), "scalar": np.array(3.5, dt)}
        path = tmp_path / f"m_{np.dtype(dt).name}.bin"
        checkpoint.write_state_dict(path, sd)
        back = checkpoint.read_state_dict(path)
        assert set(back) == set(sd)
        for k in sd:
            assert back[k].dtype == dt and back[k].shape == sd[k].shape
            np.testing.assert_array_equal(back[k], sd[k])


def test_reader_resolves_stride_and_offset(tmp_path):
    # a transposed 2x3 view with ptr_offset 1 over 7 stored values (Matrix::new(ptr, shape, stride), impl_serde.rs:160-168)
    data = [9.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0]
    img = _u64(2) + _entry("t", [2, 3], [1, 2], data, "f64", ptr_offset=1) + _entry("v", [2], [1], [7.0, 8.0], "f64")
    path = tmp_path / "s.bin"
    path.write_bytes(img)
    back = checkpoint.read_state_dict(path)
    np.testing.assert_array_equal(back["t"], np.array([[1.0, 3.0, 5.0], [2.0, 4.0, 6.0]]))
    np.testing.assert_array_equal(back["v"], np.array([7.0, 8.0]))


def test_reader_errors(tmp_path):
    good = _u64(1) + _entry("w", [2], [1], [1.0, 2.0], "f32")
    p = tmp_path / "trunc.bin"
    p.write_bytes(good[:-5])
    with pytest.raises(ZenuB200Error):
        checkpoint.read_state_dict(p)
    p = tmp_path / "type.bin"
    p.write_bytes(_u64(1) + _s("w") + _u64(1) + _u64(2) + _u64(1) + _u64(1) + _u64(2) + struct.pack("<2f", 1.0, 2.0) + _s("i32") + _u64(0))
    with pytest.raises(ZenuB200Error):
        checkpoint.read_state_dict(p)
    with pytest.raises(ZenuB200Error):
        checkpoint.read_state_dict(tmp_path / "missing.bin")
    with pytest.raises(ZenuB200Error):
        checkpoint.write_state_dict(tmp_path / "mixed.bin", {"a": np.zeros(2, np.float32), "b": np.zeros(2, np.float64)})


@pytest.mark.gpu
def test_model_save_load_reference_layout(tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from zenu_b200 import nn, ops
    ctx = ops.Context()
    a = nn.Model(ctx, "small_cnn", 10, seed=1)
    b = nn.Model(ctx, "small_cnn", 10, seed=2)
    path = tmp_path / "small_cnn.bin"
    a.save(path)
    sd = checkpoint.read_state_dict(path)
    pa = a.named_parameters()
    assert set(sd) == set(pa)
    # reference layouts in the file: filters KCRS, conv bias [1,K,1,1]
    f = pa["conv2.conv2d.filter"]["data"]
    assert sd["conv2.conv2d.filter"].shape == (64, 32, 3, 3)
    np.testing.assert_array_equal(sd["conv2.conv2d.filter"], a.filter_to_kcrs(f).cpu().numpy())
    assert sd["conv1.conv2d.bias"].shape == (1, 32, 1, 1)
    assert sd["linear1.linear.weight"].shape == (512, 64 * 32 * 32)
    b.load(path)
    pb = b.named_parameters()
    for k in pa:
        assert torch.equal(pa[k]["data"], pb[k]["data"]), k
    # a file naming a parameter the model does not have is an error (load_model returns Err); partial files are fine
    extra = dict(sd)
    extra["nope.weight"] = np.zeros((2,), np.float32)
    checkpoint.write_state_dict(tmp_path / "extra.bin", extra)
    with pytest.raises(ZenuB200Error):
        b.load(tmp_path / "extra.bin")
    checkpoint.write_state_dict(tmp_path / "partial.bin", {"linear2.linear.bias": np.full((10,), 0.25, np.float32)})
    b.load(tmp_path / "partial.bin")
    assert float(b.named_parameters()["linear2.linear.bias"]["data"].min()) == 0.25
    a.close(); b.close(); ctx.close()


def test_reader_rejects_hostile_shapes(tmp_path):
    """A corrupt file must come back as the reference's "Failed to load model" error, never as an exception escaping the C
    boundary or a giant allocation: extents are file content."""
    data = [1.0, 2.0]
    for shape, stride in (([1 << 40], [1]), ([1 << 62, 4], [4, 1]), ([(1 << 64) - 3], [1]), ([1 << 33, 1 << 33], [1, 1])):
        p = tmp_path / "hostile.bin"
        p.write_bytes(_u64(1) + _entry("w", shape, stride, data, "f32"))
        with pytest.raises(ZenuB200Error, match="Failed to load model"):
            checkpoint.read_state_dict(p)
    p = tmp_path / "count.bin"
    p.write_bytes(_u64(1 << 40))
    with pytest.raises(ZenuB200Error, match="Failed to load model"):
        checkpoint.read_state_dict(p)


@pytest.mark.gpu
@pytest.mark.parametrize("opt", ["adamw", "adam", "sgd"])
def test_resume_from_state_file_equals_uninterrupted_training(tmp_path, opt):
    """zb_model_save_state / zb_model_load_state (parameters + Adam m / v + step): 3 steps, save, restore into a
    freshly created model, 3 more steps == 6 uninterrupted steps, bit for bit (losses, parameters, BN running statistics)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from zenu_b200 import nn, ops
    rng = np.random.default_rng(12)
    xs = [torch.from_numpy(rng.standard_normal((16, 3, 32, 32)).astype(np.float32)).cuda() for _ in range(6)]
    t = np.zeros((16, 10), np.float32)
    t[np.arange(16), rng.integers(0, 10, 16)] = 1.0
    T = torch.from_numpy(t).cuda()
    kw = dict(kind=opt, lr=1e-3, weight_decay=0.01 if opt == "adamw" else 0.0)
    ctx = ops.Context()
    a = nn.Model(ctx, "small_cnn", 10, seed=1)
    a.set_optimizer(**kw)
    ref = [a.train_step(x, T, read_loss=True) for x in xs]
    b = nn.Model(ctx, "small_cnn", 10, seed=1)
    b.set_optimizer(**kw)
    first = [b.train_step(x, T, read_loss=True) for x in xs[:3]]
    path = tmp_path / "state.bin"
    b.save_state(path)
    sd = checkpoint.read_state_dict(path)
    assert float(sd["optimizer.step"]) == 3.0
    if opt != "sgd":   # Adam state is stored per parameter, in the parameter's reference layout (filters KCRS)
        assert sd["optimizer.m.conv2.conv2d.filter"].shape == (64, 32, 3, 3) and sd["optimizer.v.linear2.linear.bias"].shape == (10,)
        assert "optimizer.m.batch_norm1.batch_norm_2d.mean" not in sd
        assert np.abs(sd["optimizer.v.conv2.conv2d.filter"]).max() > 0
    else:
        assert not any(k.startswith("optimizer.m.") for k in sd)
    c = nn.Model(ctx, "small_cnn", 10, seed=99)          # different initial weights: everything must come from the file
    c.set_optimizer(**kw)
    c.load_state(path)
    rest = [c.train_step(x, T, read_loss=True) for x in xs[3:]]
    assert first + rest == ref
    pa, pc = a.named_parameters(), c.named_parameters()
    for k in pa:
        assert torch.equal(pa[k]["data"], pc[k]["data"]), k
    if opt != "sgd":   # Adam state into an SGD model is refused before anything is written
        d = nn.Model(ctx, "small_cnn", 10, seed=5)
        d.set_optimizer("sgd", lr=1e-3)
        before = d.named_parameters()["linear2.linear.weight"]["data"].clone()
        with pytest.raises(ZenuB200Error):
            d.load_state(path)
        assert torch.equal(before, d.named_parameters()["linear2.linear.weight"]["data"])
        d.close()
    # a plain model file (reference format) is a valid state file, and a state file is NOT a valid plain model file
    c.save(tmp_path / "plain.bin")
    c.load_state(tmp_path / "plain.bin")
    with pytest.raises(ZenuB200Error):
        c.load(path)
    a.close(); b.close(); c.close(); ctx.close()
