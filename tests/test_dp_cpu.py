"""Data-parallel host logic on CPU (no GPU): the bucket plan exported by the C ABI (zb_dp_plan_buckets) and, with a
world_size-2 `gloo` group, the exchange protocol bench.py / the host model use — per-rank shard gradients written into
the flat buffer at the planned offsets, one sum-allreduce per bucket in bucket order, 1/world folded into the optimizer
step — checked against single-process gradient averaging computed with the oracle."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import zenu_oracle_model as zm  # noqa: E402

KIND = {"weight": 0, "bias": 1, "buffer": 2}


def plan(arch, classes, bucket_bytes):
    from zenu_b200 import _lib
    lib = _lib.load()
    shapes = zm.param_shapes(arch, classes)
    names = list(shapes)
    n = len(names)
    numel = (ctypes.c_int64 * n)(*[int(np.prod(shapes[k][0])) for k in names])
    kind = (ctypes.c_int * n)(*[KIND[shapes[k][1]] for k in names])
    bucket = (ctypes.c_int * n)()
    offset = (ctypes.c_int64 * n)()
    nb, total, buf = ctypes.c_int(), ctypes.c_int64(), ctypes.c_int64()
    rc = lib.zb_dp_plan_buckets(numel, kind, n, bucket_bytes, 4, bucket, offset, ctypes.byref(nb), ctypes.byref(total), ctypes.byref(buf))
    assert rc == 0
    return names, list(numel), list(kind), list(bucket), list(offset), nb.value, total.value, buf.value


def test_bucket_plan_resnet50():
    names, numel, kind, bucket, offset, nb, total, buf = plan("resnet50", 1000, 25 << 20)
    trainable = [i for i in range(len(names)) if kind[i] != 2]
    assert sum(numel[i] for i in trainable) == 25557032          # ResNet-50 parameter count (25.56 M)
    assert 4 <= nb <= 6                                           # ~102 MB of gradients in ~25 MB buckets
    # every tensor 16-byte aligned, no overlap, everything inside [0, total)
    spans = sorted((offset[i], offset[i] + numel[i]) for i in trainable)
    assert all(o % 4 == 0 for o, _ in spans)
    assert all(spans[j][1] <= spans[j + 1][0] for j in range(len(spans) - 1)) and spans[-1][1] <= total
    # buckets are contiguous ranges in the flat buffer, ordered bucket 0 | bucket 1 | ...
    for b in range(nb):
        mem = [i for i in trainable if bucket[i] == b]
        assert mem and sum(numel[i] for i in mem) * 4 <= (25 << 20) + max(numel[i] for i in mem) * 4
        if b + 1 < nb:
            nxt = [i for i in trainable if bucket[i] == b + 1]
            assert max(offset[i] + numel[i] for i in mem) <= min(offset[i] for i in nxt)
        # weights before biases inside a bucket (AdamW decays the weight run only)
        w_hi = max([offset[i] + numel[i] for i in mem if kind[i] == 0], default=-1)
        b_lo = min([offset[i] for i in mem if kind[i] == 1], default=1 << 62)
        assert w_hi <= b_lo
    # reverse order: the last layer (fc) is in bucket 0, the stem in the last bucket
    assert bucket[names.index("fc.linear.weight")] == 0 and bucket[names.index("conv1.conv2d.filter")] == nb - 1
    # buffers (BN running statistics) have no gradient slot
    assert all(bucket[i] == -1 for i in range(len(names)) if kind[i] == 2) and buf > 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        arch, classes, per_rank = "small_cnn", 10, 4
        names, numel, kind, bucket, offset, nb, total, _ = plan(arch, classes, 1 << 20)   # 1 MB buckets -> several buckets
        params = zm.init_params(arch, classes, seed=42)
        rng = np.random.default_rng(7)
        x = rng.standard_normal((per_rank * world, 3, 32, 32)).astype(np.float32)
        t = np.zeros((per_rank * world, classes), np.float32)
        t[np.arange(per_rank * world), rng.integers(0, classes, per_rank * world)] = 1.0
        shard = slice(rank * per_rank, (rank + 1) * per_rank)
        model = zm.OracleModel(arch, classes, {k: v.copy() for k, v in params.items()})
        _, grads = model.forward_backward(x[shard], t[shard])
        flat = torch.zeros(total, dtype=torch.float32)
        for i, k in enumerate(names):
            if kind[i] != 2:
                flat[offset[i]:offset[i] + numel[i]] = torch.from_numpy(np.ascontiguousarray(grads[k]).ravel())
        # one allreduce per bucket, bucket 0 (closest to the loss) first: the order backward completes them in
        for b in range(nb):
            mem = [i for i in range(len(names)) if bucket[i] == b]
            lo, hi = min(offset[i] for i in mem), max(offset[i] + numel[i] for i in mem)
            dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM)
        # Optimizer::update with the 1/world gradient scale folded in (sgd.rs:20-30)
        lr = 0.01
        new = {k: (params[k].ravel() - lr * (flat[offset[i]:offset[i] + numel[i]].numpy() / world)) for i, k in enumerate(names) if kind[i] != 2}
        if rank == 0:
            np.savez(out, **new)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_bucketed_allreduce_matches_gradient_averaging(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "rank0.npz")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    # single process: average of the per-shard gradients (per-replica BatchNorm statistics, as the data-parallel path has)
    arch, classes, per_rank = "small_cnn", 10, 4
    params = zm.init_params(arch, classes, seed=42)
    rng = np.random.default_rng(7)
    x = rng.standard_normal((per_rank * world, 3, 32, 32)).astype(np.float32)
    t = np.zeros((per_rank * world, classes), np.float32)
    t[np.arange(per_rank * world), rng.integers(0, classes, per_rank * world)] = 1.0
    acc = None
    for r in range(world):
        m = zm.OracleModel(arch, classes, {k: v.copy() for k, v in params.items()})
        _, g = m.forward_backward(x[r * per_rank:(r + 1) * per_rank], t[r * per_rank:(r + 1) * per_rank])
        acc = g if acc is None else {k: acc[k] + g[k] for k in g}
    for k, g in acc.items():
        ref = params[k].ravel() - 0.01 * (g.ravel() / world)
        np.testing.assert_allclose(got[k], ref, rtol=1e-6, atol=1e-7, err_msg=k)


def test_reference_arm_other_ranks_exit_quietly():
    """bench.py --impl reference under torchrun: rank 0 alone runs and prints; the other ranks exit 0 without work."""
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
