#!/usr/bin/env python
"""Headline benchmark: clips/sec of one SlowFast-8x8-R50 training step on N B200s (BASELINE.json configs[1]).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's CPU path (oracle port) on the host cores

A step = zero_grad -> forward (engine kernels) -> cross-entropy -> backward (engine kernels) -> [one NCCL all-reduce
of the flat gradient bucket when N > 1] -> SGD step, on B=8 clips per GPU of synthetic Kinetics-shaped input
(3 x 32 x 224 x 224 fast / 3 x 8 x 224 x 224 slow, random-init weights), parity mode (split-bf16 operands, fp32
storage/accumulate).  One JSON line is printed by rank 0; see DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU_BATCH = 8          # TRAIN.BATCH_SIZE 64 / 8 GPUs (configs/Kinetics/SLOWFAST_8x8_R50.yaml)
FWD_GFLOP_PER_CLIP = 100.62  # algorithmic 2*MAC FLOPs of one forward at 224^2


def host_threads() -> int:
    """Threads the CPU legs may use: the cores this process is actually allowed to run on (cgroup / affinity), never
    more than torch's own default - asking for every core of a shared host oversubscribes and runs ~100x slower."""
    try:
        allowed = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        allowed = os.cpu_count() or 1
    return max(1, min(allowed, torch.get_num_threads(), 64))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops_burst=p["bf16_tflops"],
                    tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def reference_arm(args):
    """The reference's own CPU implementation of the path = its ATen operator sequence, restated in
    oracle/torch_oracle.py (the Python reference itself cannot travel to the GPU box; the restatement is pinned to
    it by oracle/make_golden.py).  Each step = forward + backward of a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import torch_oracle as TO
    from slowfast_b200.config import get_cfg
    torch.set_num_threads(host_threads())
    cfg = get_cfg("SLOWFAST_8x8_R50", MODEL={"DROPOUT_RATE": 0.0})
    sample_b = 2
    from slowfast_b200.nets.resnet import B200SlowFast
    torch.manual_seed(cfg.RNG_SEED)
    state = {k: v.clone() for k, v in B200SlowFast(cfg).state_dict().items()}
    inputs = TO.synthetic_inputs(cfg, sample_b, 1234)
    dlogits = torch.randn(sample_b, cfg.MODEL.NUM_CLASSES) / sample_b
    steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    for _ in range(warm):
        TO.forward_backward(cfg, state, inputs, dlogits)
    t0 = time.perf_counter()
    for _ in range(steps):
        TO.forward_backward(cfg, state, inputs, dlogits)
    dt = (time.perf_counter() - t0) / steps
    v = sample_b / dt
    line = dict(metric="clips/sec (fwd+bwd) SlowFast-8x8-R50", value=v, unit="clips/s", n_gpus=args.gpus, steps=steps,
                warmup=warm, ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload="SlowFast-8x8-R50 train step fwd+bwd, 32x224x224 fast / 8x224x224 slow",
                            per_step_sample=f"{sample_b} clips on the host CPU", threads=torch.get_num_threads()),
                cpu_baseline=dict(value=v, unit="clips/s", cores=torch.get_num_threads(), kind="port",
                                  sample=f"{steps} x fwd+bwd of {sample_b} clips (oracle/torch_oracle.py, ATen CPU fp32)"),
                e2e=dict(value=v, unit="clips/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nsplit", type=int, default=3, choices=[1, 3])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mvit", action="store_true", help="skip the secondary MViTv2-S measurement")
    ap.add_argument("--no-x3d", action="store_true", help="skip the secondary X3D-M measurement")
    ap.add_argument("--no-maskfeat", action="store_true", help="skip the secondary MaskFeat (MViTv2-S) measurement")
    ap.add_argument("--no-aten-gpu", action="store_true",
                    help="skip timing the reference's own ATen/cuDNN code path on this GPU (N=1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    import torch.nn.functional as F

    from slowfast_b200 import ops
    from slowfast_b200.config import get_cfg
    from slowfast_b200.nets.resnet import B200SlowFast

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=600))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    cfg = get_cfg("SLOWFAST_8x8_R50", B200={"NSPLIT": args.nsplit})
    torch.manual_seed(cfg.RNG_SEED)
    model = B200SlowFast(cfg).to(dev).train()
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4, nesterov=True)
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    T, A = cfg.DATA.NUM_FRAMES, cfg.SLOWFAST.ALPHA
    clip = torch.randn(B, 3, T, 224, 224, generator=g)
    idx = torch.linspace(0, T - 1, T // A).long()
    host = [clip.index_select(2, idx).contiguous().pin_memory(), clip.pin_memory()]
    labels_h = torch.randint(0, cfg.MODEL.NUM_CLASSES, (B,), generator=g).pin_memory()
    resident = [t.to(dev) for t in host]
    labels = labels_h.to(dev)

    def step(x, y):
        opt.zero_grad(set_to_none=True)
        logits = model(x)
        loss = F.cross_entropy(logits, y)
        loss.backward()
        if world > 1:
            model.allreduce_gradients()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    for _ in range(args.warmup):
        step(resident, labels)
    barrier()

    # ---- (1) device-resident throughput
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ops.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step(resident, labels)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ops.launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = B * world * args.steps / (ms_total * 1e-3)

    # ---- (2) end to end through the public call with HOST buffers: pinned H2D every step (prefetched on a copy
    #          stream, as a loader with non_blocking copies would) + D2H read of the loss every step
    copy_stream = torch.cuda.Stream()
    bufs = [[torch.empty_like(t, device=dev) for t in host] for _ in range(2)]
    lab_bufs = [torch.empty_like(labels_h, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for d, s in zip(bufs[slot], host):
                d.copy_(s, non_blocking=True)
            lab_bufs[slot].copy_(labels_h, non_blocking=True)
            ready[slot].record(copy_stream)

    for s in range(2):
        consumed[s].record()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    prefetch(0)
    last = 0.0
    for i in range(args.steps):
        slot = i & 1
        if i + 1 < args.steps:
            prefetch(slot ^ 1)
        torch.cuda.current_stream().wait_event(ready[slot])
        loss = step(bufs[slot], lab_bufs[slot])
        consumed[slot].record()
        last = loss.item()  # device -> host read of the step's result
    t1.record()
    barrier()
    ms_e2e = max_over_ranks(t0.elapsed_time(t1))
    e2e_value = B * world * args.steps / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host) + labels_h.numel() * labels_h.element_size()

    # ---- (3) per-kernel-class timing of one extra step with CUDA events around every conv launch (same stream)
    peaks = load_peaks()
    # (every rank runs the extra step - it contains the gradient all-reduce, so the collective sequence must match on
    # all ranks; rank 0's numbers are the ones reported)
    roofline = profile_conv_kernels(model, step, resident, labels, peaks, B)

    # ---- (4) CPU baseline: the oracle port on this box's host cores, bounded sample (rank 0, N == 1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_leg(cfg)

    # ---- (5) second headline model of BASELINE.json's metric: MViTv2-S 16x224x224 train step (B=4/GPU, AdamW)
    mvit = None
    if not args.no_mvit:
        try:
            mvit = mvit_leg(args, dev, world, rank, barrier, max_over_ranks)
        except Exception as e:  # noqa: BLE001 - the headline line must still be printed
            mvit = dict(error=repr(e)[:300])

    # ---- (6) third model family of the metric: X3D-M 16x224x224 train step (B=16/GPU, SGD) - HBM-bound
    x3d = None
    if not args.no_x3d:
        try:
            x3d = x3d_leg(args, dev, world, rank, barrier, max_over_ranks)
        except Exception as e:  # noqa: BLE001
            x3d = dict(error=repr(e)[:300])

    # ---- (7) MaskFeat pre-training step on the MViTv2-S encoder (masked_ssl yaml, B=4/GPU, AdamW, HOG targets inside)
    maskfeat = None
    if not args.no_maskfeat:
        try:
            maskfeat = maskfeat_leg(args, dev, world, rank, barrier, max_over_ranks)
        except Exception as e:  # noqa: BLE001
            maskfeat = dict(error=repr(e)[:300])

    # ---- (8) the reference's own GPU code path (ATen / cuDNN) on this device, N == 1 only (a reported baseline)
    aten_gpu = None
    if rank == 0 and world == 1 and not args.no_aten_gpu:
        try:
            aten_gpu = aten_gpu_leg(cfg, dev, batch=B)
        except Exception as e:  # noqa: BLE001
            aten_gpu = dict(error=repr(e)[:300])

    if rank == 0:
        step_flops = 3.0 * FWD_GFLOP_PER_CLIP * 1e9  # training step ~ 3x forward
        line = dict(
            metric="clips/sec (fwd+bwd) SlowFast-8x8-R50", value=value, unit="clips/s", n_gpus=world,
            steps=args.steps, warmup=args.warmup, ms_per_step=ms_total / args.steps, higher_is_better=True,
            scaling="weak", vs_baseline=None,
            dtype="bf16x3-split operands, f32 accumulate/storage" if args.nsplit == 3 else "bf16 operands, f32 accumulate",
            data="synthetic",
            config=dict(workload="SlowFast-8x8-R50 (configs/Kinetics/SLOWFAST_8x8_R50.yaml) train step: fwd + CE loss + "
                                 "bwd + grad all-reduce (N>1) + SGD, 32x224x224 fast / 8x224x224 slow, random init",
                        per_gpu_batch=B, global_batch=B * world, parallelism=f"dp{world}", precision_mode=f"nsplit{args.nsplit}",
                        l2_policy="per-step working set (inputs 193 MB + activations > 10 GB) exceeds the 126 MB L2; no flush needed"),
            e2e=dict(value=e2e_value, unit="clips/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                     ms_per_step=ms_e2e / args.steps, last_loss=last),
            gpu_launches=launches,
            clocks=clocks,
            roofline=roofline,
            cpu_baseline=cpu_baseline,
            mvitv2_s=mvit,
            x3d_m=x3d,
            maskfeat_s=maskfeat,
            aten_gpu_baseline=aten_gpu,
            model_tflops=dict(algorithmic_tflops=value * step_flops / 1e12,
                              frac_of_bf16_sustained=value * step_flops / 1e12 / world / peaks["tflops_sustained"],
                              peaks=peaks["source"]),
        )
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def mvit_leg(args, dev, world, rank, barrier, max_over_ranks):
    """clips/s of one MViTv2-S 16x4 train step (fwd + CE + bwd + [all-reduce] + AdamW), 4 clips per GPU
    (configs/Kinetics/MVITv2_S_16x4.yaml: BATCH_SIZE 16 x NUM_SAMPLE 2 over 8 GPUs), device-resident inputs."""
    import torch.nn.functional as F

    from slowfast_b200 import ops
    from slowfast_b200.config import get_cfg
    from slowfast_b200.nets.mvit import B200MViT
    cfg = get_cfg("MVITv2_S_16x4", B200={"NSPLIT": args.nsplit})
    torch.manual_seed(cfg.RNG_SEED)
    model = B200MViT(cfg).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5, weight_decay=0.05)
    B = 4
    g = torch.Generator().manual_seed(4321 + rank)
    x = [torch.randn(B, 3, cfg.DATA.NUM_FRAMES, 224, 224, generator=g).to(dev)]
    y = torch.randint(0, cfg.MODEL.NUM_CLASSES, (B,), generator=g).to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = F.cross_entropy(model(x), y)
        loss.backward()
        if world > 1:
            model.allreduce_gradients()
        opt.step()
        return loss

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    l0 = ops.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    return dict(metric="clips/sec (fwd+bwd) MViTv2-S", value=B * world * args.steps / (ms * 1e-3), unit="clips/s",
                ms_per_step=ms / args.steps, per_gpu_batch=B, gpu_launches=ops.launches() - l0,
                algorithmic_tflops=B * world * args.steps / (ms * 1e-3) * 3 * 128.45e9 / 1e12,
                config="configs/Kinetics/MVITv2_S_16x4.yaml, drop-path 0.2 + head dropout 0.5 on, AdamW, synthetic",
                last_loss=float(loss.item()))


def x3d_leg(args, dev, world, rank, barrier, max_over_ranks):
    """clips/s of one X3D-M train step (fwd + CE + bwd + [all-reduce] + SGD-nesterov), 16 clips per GPU
    (configs/Kinetics/X3D_M.yaml: BATCH_SIZE 128 over 8 GPUs), device-resident inputs.  The model is HBM-bound
    (365.6 MB ideal forward traffic per clip, ~3x that for a train step)."""
    import torch.nn.functional as F

    from slowfast_b200 import ops
    from slowfast_b200.config import get_cfg
    from slowfast_b200.nets.x3d import B200X3D
    cfg = get_cfg("X3D_M", B200={"NSPLIT": args.nsplit})
    torch.manual_seed(cfg.RNG_SEED)
    model = B200X3D(cfg).to(dev).train()
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, nesterov=True, weight_decay=5e-5)
    B = 16
    g = torch.Generator().manual_seed(5321 + rank)
    x = [torch.randn(B, 3, cfg.DATA.NUM_FRAMES, 224, 224, generator=g).to(dev)]
    y = torch.randint(0, cfg.MODEL.NUM_CLASSES, (B,), generator=g).to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = F.cross_entropy(model(x), y)
        loss.backward()
        if world > 1:
            model.allreduce_gradients()
        opt.step()
        return loss

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    l0 = ops.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    cps = B * world * args.steps / (ms * 1e-3)
    return dict(metric="clips/sec (fwd+bwd) X3D-M", value=cps, unit="clips/s", ms_per_step=ms / args.steps,
                per_gpu_batch=B, gpu_launches=ops.launches() - l0,
                algorithmic_tflops=cps * 3 * 9.47e9 / 1e12,
                ideal_traffic_gbps=cps * 3 * 365.6e6 / 1e9,
                config="configs/Kinetics/X3D_M.yaml, head dropout 0.5 on, SGD-nesterov, synthetic",
                last_loss=float(loss.item()))


def maskfeat_leg(args, dev, world, rank, barrier, max_over_ranks):
    """clips/s of one MaskFeat pre-training step (configs/masked_ssl/k400_MVITv2_S_16x4_MaskFeat_PT.yaml): mask-token
    MViT encoder fwd + HOG targets + MultipleMSELoss + bwd + [all-reduce] + AdamW, 4 clips per GPU (BATCH_SIZE 32 / 8),
    40 % of the 8x7x7 cube cells masked, device-resident inputs."""
    import torch.nn.functional as F

    from slowfast_b200 import ops
    from slowfast_b200.config import get_cfg
    from slowfast_b200.nets.maskfeat import B200MaskMViT
    cfg = get_cfg("MVITv2_S_16x4_MaskFeat_PT", B200={"NSPLIT": args.nsplit})
    torch.manual_seed(cfg.RNG_SEED)
    model = B200MaskMViT(cfg).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5, weight_decay=0.05)
    B = 4
    g = torch.Generator().manual_seed(6321 + rank)
    frames = torch.randn(B, 3, cfg.DATA.NUM_FRAMES, 224, 224, generator=g).to(dev)
    mask = (torch.rand(B, 8, 7, 7, generator=g) < 0.4).float().to(dev)
    meta = torch.Tensor()

    def step():
        opt.zero_grad(set_to_none=True)
        preds, labels = model([frames, meta, mask])
        loss = sum(F.mse_loss(p, l[0]) * l[1] for p, l in zip(preds, labels))  # losses.py:38-62 MultipleMSELoss
        loss.backward()
        if world > 1:
            model.allreduce_gradients()
        opt.step()
        return loss

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    l0 = ops.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    cps = B * world * args.steps / (ms * 1e-3)
    return dict(metric="clips/sec (fwd+bwd) MaskFeat MViTv2-S", value=cps, unit="clips/s", ms_per_step=ms / args.steps,
                per_gpu_batch=B, gpu_launches=ops.launches() - l0, algorithmic_tflops=cps * 3 * 173.0e9 / 1e12,
                config="configs/masked_ssl/k400_MVITv2_S_16x4_MaskFeat_PT.yaml, 40 % cube mask, AdamW, synthetic",
                last_loss=float(loss.item()))


def profile_conv_kernels(model, step, resident, labels, peaks, B):
    """Time every implicit-GEMM launch of one step with CUDA events on the launching stream and aggregate per kernel
    class; the roofline object describes the class with the largest share of the step."""
    from slowfast_b200 import ops
    recs = []
    orig_conv, orig_wgrad = ops.conv_igemm, ops.conv_wgrad

    def timed(kind, fn, flops_fn, bytes_fn):
        def wrapper(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **k)
            e.record()
            recs.append((kind, s, e, flops_fn(*a, **k), bytes_fn(*a, **k)))
            return r
        return wrapper

    def conv_flops(x, f, geom, out, strides, **k):
        m = x.n * geom.out[0] * geom.out[1] * geom.out[2]
        return 2.0 * m * f.rows * f.ntaps * f.cols_pad

    def conv_bytes(x, f, geom, out, strides, **k):
        m = x.n * geom.out[0] * geom.out[1] * geom.out[2]
        planes = 2 if x.lo is not None else 1
        return x.rows * x.c * 2 * planes + f.rows * f.ntaps * f.cols_pad * 2 * planes + m * f.rows * 4

    def wg_flops(x, dy, geom, dwm, **k):
        taps = geom.k[0] * geom.k[1] * geom.k[2]
        return 2.0 * dy.rows * dy.c * taps * x.c

    def wg_bytes(x, dy, geom, dwm, **k):
        planes = 2 if x.lo is not None else 1
        return (x.rows * x.c + dy.rows * dy.c) * 2 * planes + dwm.numel() * 4

    import slowfast_b200.engine as eng
    ops.conv_igemm = timed("conv_igemm(fprop+dgrad)", orig_conv, conv_flops, conv_bytes)
    ops.conv_wgrad = timed("conv_wgrad", orig_wgrad, wg_flops, wg_bytes)
    graphs_were = model.cuda_graphs
    model.cuda_graphs = False  # replays bypass the python wrappers: time the same launches eagerly
    try:
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        step(resident, labels)
        s1.record()
        torch.cuda.synchronize()
    finally:
        ops.conv_igemm, ops.conv_wgrad = orig_conv, orig_wgrad
        model.cuda_graphs = graphs_were
    step_ms = s0.elapsed_time(s1)
    agg = {}
    for kind, s, e, fl, by in recs:
        a = agg.setdefault(kind, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
        a["ms"] += s.elapsed_time(e)
        a["flops"] += fl
        a["bytes"] += by
        a["launches"] += 1
    if not agg:
        return None
    top = max(agg, key=lambda k: agg[k]["ms"])
    a = agg[top]
    tf = a["flops"] / (a["ms"] * 1e-3) / 1e12
    gbs = a["bytes"] / (a["ms"] * 1e-3) / 1e9
    tensor_frac = tf / peaks["tflops_sustained"]
    hbm_frac = gbs / peaks["hbm_gbs"]
    bound = "tensor" if tensor_frac >= hbm_frac else "hbm"
    return dict(kernel=top, bound=bound,
                achieved=tf if bound == "tensor" else gbs, peak=peaks["tflops_sustained"] if bound == "tensor" else peaks["hbm_gbs"],
                unit="TFLOP/s" if bound == "tensor" else "GB/s", frac=max(tensor_frac, hbm_frac), traffic=None,
                peak_source=peaks["source"] + (" (sustained bf16: kernel timed inside a long step)" if bound == "tensor" else ""),
                per_class={k: dict(launches=v["launches"], ms=round(v["ms"], 3), share_of_step=round(v["ms"] / step_ms, 4),
                                   algorithmic_tflops=round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2),
                                   algorithmic_gbs=round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)) for k, v in agg.items()},
                profiled_step_ms=round(step_ms, 3),
                note="achieved = sum of algorithmic 2*M*N*K FLOPs (operand-split passes NOT counted) / sum of CUDA-event "
                     "launch durations of the class in one step; events add launch gaps, so this is a lower bound")


def cpu_baseline_leg(cfg):
    from oracle import torch_oracle as TO
    from slowfast_b200.nets.resnet import B200SlowFast
    torch.set_num_threads(host_threads())
    c = cfg.clone()
    c.MODEL.DROPOUT_RATE = 0.0
    torch.manual_seed(cfg.RNG_SEED)
    state = {k: v.clone() for k, v in B200SlowFast(c).state_dict().items()}
    b = 2
    inputs = TO.synthetic_inputs(c, b, 1234)
    dlogits = torch.randn(b, c.MODEL.NUM_CLASSES) / b
    TO.forward_backward(c, state, inputs, dlogits)  # warm-up
    n, t0 = 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 20):
        TO.forward_backward(c, state, inputs, dlogits)
        n += 1
    dt = (time.perf_counter() - t0) / n
    return dict(value=b / dt, unit="clips/s", cores=torch.get_num_threads(), kind="port",
                sample=f"{n} x fwd+bwd of {b} clips, fp32 ATen CPU kernels (oracle/torch_oracle.py restatement of the "
                       f"reference's nn.Conv3d/BatchNorm3d path), {dt * 1e3:.0f} ms/iter")


def aten_gpu_leg(cfg, dev, batch: int = 8, iters: int = 3):
    """The comparator SURVEY.md section 8(d) asks for next to the CPU baseline: the reference's OWN operator sequence
    (nn.Conv3d / BatchNorm3d / ... = ATen + cuDNN kernels, restated in oracle/torch_oracle.py) timed on the SAME
    device, fwd + bwd of the same SlowFast batch, in fp32 (TF32 off: the reference's parity setting), with TF32
    allowed, and under bf16 autocast.  A reported baseline, never on the product path."""
    from oracle import torch_oracle as TO
    from slowfast_b200.nets.resnet import B200SlowFast
    c = cfg.clone()
    c.MODEL.DROPOUT_RATE = 0.0
    torch.manual_seed(cfg.RNG_SEED)
    state = {k: v.to(dev) for k, v in B200SlowFast(c).state_dict().items()}
    inputs = [t.to(dev) for t in TO.synthetic_inputs(c, batch, 1234)]
    dlogits = (torch.randn(batch, c.MODEL.NUM_CLASSES) / batch).to(dev)
    is_cuda = torch.device(dev).type == "cuda"
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    out = {}
    try:
        torch.backends.cudnn.benchmark = True
        for mode in ("fp32", "tf32", "bf16_autocast"):
            torch.backends.cudnn.allow_tf32 = mode != "fp32"
            torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"

            def one():
                if mode == "bf16_autocast":
                    with torch.autocast(torch.device(dev).type, dtype=torch.bfloat16):
                        TO.forward_backward(c, state, inputs, dlogits)
                else:
                    TO.forward_backward(c, state, inputs, dlogits)

            for _ in range(2):
                one()
            if is_cuda:
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    one()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
            else:
                t0 = time.perf_counter()
                for _ in range(iters):
                    one()
                ms = (time.perf_counter() - t0) / iters * 1e3
            out[mode] = dict(ms_per_step=ms, clips_per_s=batch / (ms * 1e-3))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    out["what"] = (f"fwd+bwd (no optimizer step) of {batch} SlowFast-8x8-R50 clips through torch's own ATen/cuDNN kernels "
                   "on this GPU: the reference's GPU code path")
    return out


if __name__ == "__main__":
    main()
