#!/usr/bin/env python3
"""bench.py — headline benchmark: ResNet-50 f32/TF32 training throughput (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            our arm (libzenu_b200.so through the host model API)
  python bench.py --impl reference ...                      the reference's CPU path (oracle port) on the host cores

A step = one full train step (forward, cross-entropy, backward, SGD update; with N > 1 the bucketed NCCL gradient
allreduce overlapped with backward) of ResNet-50 on a synthetic batch of 256 images (3x224x224, 1000 classes) per GPU.
`value` is timed with the batch already resident in HBM; `e2e` is the same step fed from pinned HOST memory through the
library's input staging (zb_input_stage_*: the uint8 batch + int32 labels cross PCIe every step on a copy stream, are expanded on
the device, and the loss is read back to the host) — both through the public API.  `--e2e-input f32` ships the f32 batch instead.
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.  L2: every step
streams > 20 GB of activations, far larger than the 126 MB L2, so no explicit flush is needed.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "resnet50_train_images_per_s"   # BASELINE.json metric; other --arch values report <arch>_train_images_per_s
UNIT = "images/s"
FWD_GFLOP_PER_IMG = {"resnet50": 8.174 + 0.0041, "resnet18": 3.627 + 0.001, "small_cnn": 0.107}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_burst": float(p["bf16_tflops"]),
                "bf16_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    except Exception:  # noqa: BLE001  (profiling recipe's stated fallback)
        return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


TENSOR_KERNELS = ("umma_kernel", "halo_conv_kernel", "wgrad_halo_kernel", "stem_dgrad_kernel")   # every PROF_TENSOR launch


def ncu_traffic(kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (average over the launches of one step) of the kernels whose
    name starts with one of kernel_prefix, from the committed ncu summary of the same workload (profiles/*_traffic.json)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None, None
    try:
        with open(files[-1]) as f:
            t = json.load(f)
        n = b = 0.0
        for name, k in t["kernels"].items():
            if name.startswith(tuple(kernel_prefix) if not isinstance(kernel_prefix, str) else kernel_prefix) and k.get("dram_bytes_per_launch", 0) > 0:
                n += k["launches"]
                b += k["dram_bytes_per_launch"] * k["launches"]
        return (b / n if n else None), os.path.basename(files[-1])
    except Exception:  # noqa: BLE001
        return None, None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}


def synthetic_batch(batch, hw, classes, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((batch, 3, hw, hw), generator=g, dtype=torch.float32)
    labels = torch.randint(0, classes, (batch,), generator=g)
    t = torch.zeros((batch, classes), dtype=torch.float32)
    t[torch.arange(batch), labels] = 1.0
    return x, t


def cpu_reference_steps(arch, classes, hw, sample_batch, steps, warmup):
    """The reference's CPU algorithm (oracle port: im2col + OpenBLAS conv, CPU BatchNorm, ...) on the host cores."""
    import numpy as np

    from oracle import zenu_oracle as zo
    from oracle import zenu_oracle_model as zm
    cores = os.cpu_count() or 1
    blas = zo.use_openblas(threads=cores)
    model = zm.OracleModel(arch, classes, zm.init_params(arch, classes, seed=42))
    rng = np.random.default_rng(1234)
    x = rng.standard_normal((sample_batch, 3, hw, hw)).astype(np.float32)
    t = np.zeros((sample_batch, classes), np.float32)
    t[np.arange(sample_batch), rng.integers(0, classes, sample_batch)] = 1.0
    for _ in range(warmup):
        model.train_step(x, t, kind="sgd", lr=0.01)
    t0 = time.perf_counter()
    loss = 0.0
    for _ in range(steps):
        loss = model.train_step(x, t, kind="sgd", lr=0.01)
    dt = time.perf_counter() - t0
    return {"images_per_s": sample_batch * steps / dt, "sec_per_step": dt / max(steps, 1), "cores": cores,
            "blas": "OpenBLAS (numpy bundled, all cores)" if blas else "plain C loops", "loss": float(loss)}


def base_config(args, world):
    """The workload description both arms print (identical dicts: the driver compares them)."""
    return {"workload": workload_name(args), "global_batch": args.batch * world, "parallelism": f"dp{world}",
            "l2": "inputs larger than L2 (each step streams > 20 GB of activations; L2 is 126 MB)" if args.arch != "small_cnn"
                  else "L2 flushed by the step itself: 1.7 GB of activations + 134 MB of Linear weights per step vs 126 MB of L2",
            "optimizer": "SGD lr 0.01 (zenu-optimizer/src/sgd.rs)",
            "grad_allreduce": "bucketed NCCL sum, overlapped with backward" if world > 1 else "none (1 GPU)",
            "input_grad_of_conv1": "computed (as the reference does)"}


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (oracle port: no Rust toolchain here) on the host cores, every
    step a bounded sample of the workload: one full train step on `--ref-batch` of the batch's images (the whole batch for the
    CPU-runnable config, small_cnn)."""
    if rank != 0:
        return
    sample = min(args.ref_batch, args.batch) if args.arch != "small_cnn" else args.batch
    r = cpu_reference_steps(args.arch, args.classes, args.hw, sample, args.steps, args.warmup)
    ms = r["sec_per_step"] * 1e3
    line = {
        "impl": "reference", "metric": metric_name(args), "value": r["images_per_s"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": base_config(args, world),
        "cpu_baseline": {"value": r["images_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": f"{args.steps} full train steps on {sample} of the {args.batch} images of a batch after {args.warmup} warm-up steps, "
                                   f"{r['blas']}; reference CPU path restated in C (oracle/)"},
        "e2e": {"value": r["images_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def metric_name(args):
    return METRIC if args.arch == "resnet50" else f"{args.arch}_train_images_per_s"


def workload_name(args):
    return (f"{args.arch} train step (fwd + cross-entropy + bwd + SGD lr 0.01), batch {args.batch}/GPU, 3x{args.hw}x{args.hw}, "
            f"{args.classes} classes, f32 storage, TF32 tensor-core math")


def run_ours(args, rank, world, local_rank):
    import torch

    from zenu_b200 import nn, ops
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    use_graph = args.graph
    if use_graph:   # stream capture needs a real stream: everything (torch copies, events, our kernels) moves to one side stream
        torch.cuda.set_stream(torch.cuda.Stream())
    ctx = ops.Context(device=local_rank)
    lib = ctx.lib
    if world > 1:
        import ctypes
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_ubyte * 128)()
            ops.check(lib.zb_dp_unique_id(ctx.handle, buf))
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().tolist())
        ops.check(lib.zb_dp_init(ctx.handle, raw, rank, world))
    model = nn.Model(ctx, args.arch, args.classes, fused=True, seed=42, bucket_mb=args.bucket_mb)
    model.set_optimizer("sgd", lr=0.01)
    if use_graph:
        model.set_graph(True)
    x_host, t_host = synthetic_batch(args.batch, args.hw, args.classes, 1234 + rank)
    x_pin, t_pin = x_host.pin_memory(), t_host.pin_memory()
    X, T = x_pin.cuda(non_blocking=True), t_pin.cuda(non_blocking=True)
    loss_dev = torch.zeros(1, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        model.train_step(X, T, loss_out=loss_dev)
    ctx.check()
    # ---------------------------------------------------------------- timed region: inputs resident in HBM
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    import ctypes
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        model.train_step(X, T, loss_out=loss_dev)
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    value_graphs = model.graph_count()
    # same K steps again with a CUDA-event pair around every tensor-core launch / BatchNorm op (the live roofline numbers);
    # kept out of the headline region because ~280 extra event records per step cost ~2 % of it
    ops.check(lib.zb_ctx_profile_enable(ctx.handle, 1))
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pv0.record()
    for _ in range(args.steps):
        model.train_step(X, T, loss_out=loss_dev)
    pv1.record()
    torch.cuda.synchronize()
    prof_elapsed_ms = pv0.elapsed_time(pv1)
    prof = {}
    for cls, name in ((0, "tensor"), (1, "bn")):
        n_ops, ms, work = ctypes.c_int64(), ctypes.c_double(), ctypes.c_double()
        ops.check(lib.zb_ctx_profile_read(ctx.handle, cls, ctypes.byref(n_ops), ctypes.byref(ms), ctypes.byref(work)))
        prof[name] = (n_ops.value, ms.value, work.value)
    ops.check(lib.zb_ctx_profile_enable(ctx.handle, 0))
    clocks = sampler.stop() if rank == 0 else None
    final_loss = float(loss_dev.item())
    ctx.check()
    # ---------------------------------------------------------------- per-node table (outside the timed region): 2 more steps
    node_summary = None
    node_rows = None
    try:
        model.profile(True)
        for _ in range(2):
            model.train_step(X, T, loss_out=loss_dev)
        rows = model.profile_table()
        node_rows = rows
        model.profile(False)
        cls = {}
        for k, n, ms, fl, by in rows:
            c = k.split(".")[0] if k.split(".")[0] in ("conv", "bn") else "other"
            e = cls.setdefault(c, {"ms_per_step": 0.0, "flops_per_step": 0.0, "bytes_per_step": 0.0})
            e["ms_per_step"] += ms / 2; e["flops_per_step"] += fl / 2; e["bytes_per_step"] += by / 2
        node_summary = cls
    except Exception as e:  # noqa: BLE001
        node_summary = None
    # ---------------------------------------------------------------- e2e: batch comes from pinned host memory each step
    if args.e2e_input == "u8":
        # the library's input staging (zb_input_stage_*): uint8 NHWC images + int32 labels in pinned host memory -> copy stream ->
        # device expansion (normalise, NCHW, one-hot) -> train step; the copy of batch i+1 overlaps the step of batch i
        import numpy as np
        stage = nn.InputStage(ctx, args.batch, 3, args.hw, args.hw, args.classes, mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225], slots=2)
        rng = np.random.default_rng(4321 + rank)
        for sl in range(2):
            img, lab = stage.host_buffers(sl)
            img[...] = rng.integers(0, 256, img.shape, dtype=np.uint8)
            lab[...] = rng.integers(0, args.classes, lab.shape, dtype=np.int32)
        h2d_bytes = stage.h2d_bytes
        prefetch = stage.submit

        def fetch(i):
            return stage.wait(i % 2)

        def done(i):
            pass
    else:
        copy_stream = torch.cuda.Stream()
        bufs = [(torch.empty_like(X), torch.empty_like(T)) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        h2d_bytes = int(x_pin.numel() * 4 + t_pin.numel() * 4)
        for e in consumed:
            e.record()

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i % 2])
                bufs[i % 2][0].copy_(x_pin, non_blocking=True)
                bufs[i % 2][1].copy_(t_pin, non_blocking=True)
                ready[i % 2].record(copy_stream)

        def fetch(i):
            torch.cuda.current_stream().wait_event(ready[i % 2])
            return bufs[i % 2]

        def done(i):
            consumed[i % 2].record()

    # untimed: two passes over each staging buffer (with graph replay a new buffer address is a new capture after its own warm-up)
    for i in range(6):
        prefetch(i % 2)
        xb, tb = fetch(i)
        model.train_step(xb, tb, loss_out=loss_dev, read_loss=True)
        done(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prefetch(0)
    loss_host = 0.0
    # Every step's loss is read on the host inside the timed region, one step behind (zb_model_train_step_async / zb_model_loss_wait:
    # the loss lands in a pinned ring slot behind its step; the host blocks on step i-1 after it has enqueued step i, the last loss
    # is waited for before the region ends), so the device has the next step queued while the host stages the one after.
    losses_read = 0
    for i in range(args.steps):
        if i + 1 < args.steps:
            prefetch((i + 1) % 2)
        xb, tb = fetch(i)
        model.train_step_async(xb, tb, loss_dev)
        done(i)
        if i > 0:
            loss_host = model.loss_wait(1)   # D2H read of step i-1's loss
            losses_read += 1
    loss_host = model.loss_wait(0)           # ... and of the last step's
    losses_read += 1
    assert losses_read == args.steps
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    e2e_graphs = model.graph_count()
    # ---------------------------------------------------------------- reduce over ranks (max time)
    times = torch.tensor([elapsed_ms, e2e_ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = float(times[0]), float(times[1])
    global_batch = args.batch * world
    value = global_batch * args.steps / (elapsed_ms / 1e3)
    e2e_value = global_batch * args.steps / (e2e_ms / 1e3)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    tf32_peak = peaks["bf16_sustained"] / 2.0   # dense TF32 = 1/2 of the measured (sustained, in-step) bf16 GEMM rate
    t_ops, t_ms, t_flops = prof["tensor"]
    b_ops, b_ms, b_bytes = prof["bn"]
    step_ms = elapsed_ms / args.steps
    tensor_share = t_ms / prof_elapsed_ms if prof_elapsed_ms > 0 else 0.0
    bn_share = b_ms / prof_elapsed_ms if prof_elapsed_ms > 0 else 0.0
    tensor_tflops = (t_flops / (t_ms * 1e-3)) / 1e12 if t_ms > 0 else 0.0
    bn_gbs = (b_bytes / (b_ms * 1e-3)) / 1e9 if b_ms > 0 else 0.0
    if tensor_share >= bn_share:
        # the committed ncu capture is of the headline workload only (tools/ncu_step.py: ResNet-50, batch 256, 3x224x224)
        headline = args.arch == "resnet50" and args.batch == 256 and args.hw == 224
        traffic, traffic_src = ncu_traffic(TENSOR_KERNELS) if headline else (None, None)
        roofline = {"kernel": "tcgen05 kind::tf32 conv/GEMM kernels (zb::umma_kernel, halo_conv_kernel, wgrad_halo_kernel, stem_dgrad_kernel)",
                    "bound": "tensor", "achieved": tensor_tflops,
                    "peak": tf32_peak, "unit": "TFLOP/s", "frac": tensor_tflops / tf32_peak if tf32_peak else None, "traffic": traffic,
                    "traffic_source": f"profiles/{traffic_src}: ncu dram bytes per launch, averaged over the step's launches" if traffic else None,
                    "algorithmic_bytes_per_launch_avg": node_summary["conv"]["bytes_per_step"] / max(t_ops / args.steps, 1) if node_summary else None,
                    "peak_source": f"{peaks['source']}: bf16_tflops_sustained/2 (dense TF32 = half the bf16 rate)",
                    "launches_per_step": t_ops / args.steps, "share_of_step": tensor_share,
                    "algorithmic_gflop_per_launch_avg": t_flops / max(t_ops, 1) / 1e9, "avg_launch_ms": t_ms / max(t_ops, 1)}
    else:
        roofline = {"kernel": "BatchNorm2d reduce/apply kernels (fused ReLU / residual)", "bound": "hbm", "achieved": bn_gbs,
                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": bn_gbs / peaks["hbm_gbs"], "traffic": None,
                    "peak_source": f"{peaks['source']}: hbm_gbs", "ops_per_step": b_ops / args.steps, "share_of_step": bn_share,
                    "algorithmic_mb_per_op_avg": b_bytes / max(b_ops, 1) / 1e6, "avg_op_ms": b_ms / max(b_ops, 1)}
    by_node = None
    if node_summary:
        by_node = {c: {"ms_per_step": round(e["ms_per_step"], 3),
                       "achieved_tflops": round(e["flops_per_step"] / e["ms_per_step"] / 1e9, 1) if e["flops_per_step"] else None,
                       "achieved_gbs": round(e["bytes_per_step"] / e["ms_per_step"] / 1e6, 0) if e["bytes_per_step"] else None,
                       "frac_of_hbm_peak": round(e["bytes_per_step"] / e["ms_per_step"] / 1e6 / peaks["hbm_gbs"], 3) if e["bytes_per_step"] else None}
                   for c, e in node_summary.items()}
    secondary = {"tensor": {"achieved_tflops": tensor_tflops, "frac_of_tf32_peak": tensor_tflops / tf32_peak if tf32_peak else None,
                            "share_of_step": tensor_share, "launches_per_step": t_ops / args.steps},
                 "bn_hbm": {"achieved_gbs": bn_gbs, "frac_of_hbm_peak": bn_gbs / peaks["hbm_gbs"], "share_of_step": bn_share,
                            "ops_per_step": b_ops / args.steps}}
    train_tflop_per_step = 3.0 * FWD_GFLOP_PER_IMG.get(args.arch, 0.0) * args.batch / 1e3
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        full = args.arch == "small_cnn"        # BASELINE configs[0] is the CPU-runnable case: its train step is timed in full
        cb = args.batch if full else min(args.cpu_batch, args.batch)
        cs = 10 if full else 2
        r = cpu_reference_steps(args.arch, args.classes, args.hw, cb, cs, 1)
        cpu_baseline = {"value": r["images_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                        "sample": f"{cs} train steps at batch {cb} (of {args.batch}) after 1 warm-up, {r['blas']}"}
    # whole-step roofline: images / sum over tape nodes of max(flops / tensor peak, algorithmic bytes / HBM bandwidth)
    whole = None
    if node_rows:
        ideal_ms = sum(max(fl / 2 / (tf32_peak * 1e12), by / 2 / (peaks["hbm_gbs"] * 1e9)) for _, _, _, fl, by in node_rows) * 1e3
        whole = {"ideal_ms_per_step": ideal_ms, "measured_ms_per_step": step_ms, "frac": ideal_ms / step_ms if step_ms > 0 else None,
                 "ideal_images_per_s_per_gpu": args.batch / (ideal_ms / 1e3) if ideal_ms > 0 else None,
                 "definition": "sum over the step's tape nodes (forward and backward) of max(algorithmic FLOPs / dense TF32 peak, algorithmic "
                               "bytes / HBM copy bandwidth), peaks from MEASURED_PEAKS.json (sustained bf16 / 2, hbm_gbs)"}
    line = {
        "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32",
        "data": "synthetic",
        "config": base_config(args, world),
        "step_graphs": value_graphs,   # > 0: the timed steps were replayed from CUDA graphs captured during the warm-up
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps, "loss": loss_host,
                "input": "uint8 NHWC images + int32 labels through zb_input_stage_* (pinned, double-buffered, expanded on the device)"
                         if args.e2e_input == "u8" else "f32 NCHW batch + f32 one-hot targets from pinned memory (torch copy stream)",
                "loss_read": "every step's loss is copied to the host inside the timed region, read one step behind "
                             "(zb_model_train_step_async / zb_model_loss_wait); the last one is waited for before the region ends",
                "step_graphs": e2e_graphs},
        "roofline": roofline, "roofline_by_class": secondary, "tape_nodes_by_class": by_node, "whole_step_roofline": whole,
        "step_model_flops": {"algorithmic_tflop_per_step_per_gpu": train_tflop_per_step,
                             "achieved_tflops_whole_step": train_tflop_per_step / (step_ms / 1e3) if step_ms > 0 else None},
        "cpu_baseline": cpu_baseline, "final_loss": final_loss, "hbm_bytes_reserved": model.bytes_reserved(),
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arch", default="resnet50", choices=["resnet50", "resnet18", "small_cnn"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default: 256; small_cnn 64 = BASELINE configs[0])")
    ap.add_argument("--hw", type=int, default=None, help="input height = width (default: 224; small_cnn 32)")
    ap.add_argument("--classes", type=int, default=None, help="default: 1000; small_cnn 10")
    ap.add_argument("--e2e-input", default="u8", choices=["u8", "f32"], help="what the e2e leg ships across PCIe every step")
    ap.add_argument("--bucket-mb", type=int, default=25)
    ap.add_argument("--ref-batch", type=int, default=8, help="images per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--cpu-batch", type=int, default=16, help="images per step of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="run every step eagerly instead of replaying it from a CUDA graph (zb_model_set_graph; same kernels, bit-identical results)")
    args = ap.parse_args()
    small = args.arch == "small_cnn"
    args.batch = args.batch if args.batch is not None else (64 if small else 256)
    args.hw = args.hw if args.hw is not None else (32 if small else 224)
    args.classes = args.classes if args.classes is not None else (10 if small else 1000)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run (one rank per GPU)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
