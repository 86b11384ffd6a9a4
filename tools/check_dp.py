#!/usr/bin/env python3
"""Data-parallel correctness on real GPUs (run under torchrun, one rank per GPU):
after one DP train step on different shards, (1) every rank holds bit-identical parameters and (2) they equal
p0 - lr * mean_over_ranks(grad_r), where grad_r comes from a non-DP model on the same shard.  Prints one JSON line on rank 0."""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenu_b200 import nn, ops  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    arch, classes, n, hw, lr = "resnet18", 10, 8, 64, 0.05
    g = torch.Generator().manual_seed(100 + rank)           # a different shard per rank
    X = torch.randn((n, 3, hw, hw), generator=g).cuda()
    T = torch.zeros((n, classes))
    T[torch.arange(n), torch.randint(0, classes, (n,), generator=g)] = 1.0
    T = T.cuda()
    # ---- non-DP model: local gradient and the starting parameters
    ctx0 = ops.Context(device=local)
    m0 = nn.Model(ctx0, arch, classes, seed=42, bucket_mb=1)
    m0.set_optimizer("sgd", lr=lr)
    m0.forward_backward(X, T)
    ctx0.check()
    named0 = m0.named_parameters()
    names = [k for k, v in named0.items() if v["grad"] is not None]
    p0 = {k: named0[k]["data"].clone() for k in names}
    g_local = {k: named0[k]["grad"].clone() for k in names}
    # ---- DP model (same seed => same init on every rank)
    ctx = ops.Context(device=local)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (ctypes.c_ubyte * 128)()
        ops.check(ctx.lib.zb_dp_unique_id(ctx.handle, buf))
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    ops.check(ctx.lib.zb_dp_init(ctx.handle, bytes(uid.cpu().tolist()), rank, world))
    m = nn.Model(ctx, arch, classes, seed=42, bucket_mb=1)   # 1 MB buckets: many bucket allreduces overlapped with backward
    m.set_optimizer("sgd", lr=lr)
    m.train_step(X, T)
    ctx.check()
    named = m.named_parameters()
    worst_mean, worst_sync = 0.0, 0.0
    for k in names:
        gsum = g_local[k].clone()
        dist.all_reduce(gsum)
        expect = p0[k] - lr * (gsum / world)
        got = named[k]["data"]
        denom = float(expect.abs().max()) + 1e-12
        worst_mean = max(worst_mean, float((got - expect).abs().max()) / denom)
        other = got.clone()
        dist.broadcast(other, 0)
        worst_sync = max(worst_sync, float((got - other).abs().max()))
    # ---- (3) the DP step replayed from a CUDA graph (bucket allreduces captured on the comm stream) == the eager DP step
    def dp_ctx():
        c = ops.Context(device=local)
        u = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            b = (ctypes.c_ubyte * 128)()
            ops.check(c.lib.zb_dp_unique_id(c.handle, b))
            u = torch.tensor(list(b), dtype=torch.uint8)
        u = u.cuda()
        dist.broadcast(u, 0)
        torch.cuda.current_stream().synchronize()
        ops.check(c.lib.zb_dp_init(c.handle, bytes(u.cpu().tolist()), rank, world))
        return c

    finals, graphs = [], 0
    side = torch.cuda.Stream()
    for use_graph in (False, True):
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            c = dp_ctx()
            mm = nn.Model(c, arch, classes, seed=42, bucket_mb=1)
            mm.set_optimizer("sgd", lr=lr)
            if use_graph:
                mm.set_graph(True)
            lb = torch.zeros(1, device="cuda")
            for _ in range(6):
                mm.train_step(X, T, loss_out=lb)
            c.check()
            if use_graph:
                graphs = mm.graph_count()
            finals.append({k: v["data"].clone() for k, v in mm.named_parameters().items() if v["grad"] is not None})
            mm.close(); c.close()
        torch.cuda.synchronize()
    worst_graph = 0.0
    for k in finals[0]:
        denom = float(finals[0][k].abs().max()) + 1e-12
        worst_graph = max(worst_graph, float((finals[0][k] - finals[1][k]).abs().max()) / denom)
        other = finals[1][k].clone()
        dist.broadcast(other, 0)
        worst_sync = max(worst_sync, float((finals[1][k] - other).abs().max()))
    stats = torch.tensor([worst_mean, worst_sync, worst_graph, float(graphs)], device="cuda")
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    if rank == 0:
        ok = float(stats[0]) < 1e-5 and float(stats[1]) == 0.0 and float(stats[2]) < 1e-5 and int(stats[3]) == 1
        print(json.dumps({"world": world, "params_vs_mean_gradient_max_rel": float(stats[0]), "rank_divergence_max_abs": float(stats[1]),
                          "graph_vs_eager_max_rel": float(stats[2]), "step_graphs": int(stats[3]), "ok": ok}), flush=True)
    m.close(); ctx.close(); m0.close(); ctx0.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
