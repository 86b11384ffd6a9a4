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
