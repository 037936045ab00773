#!/usr/bin/env python
"""Benchmark of the B200-native FaceEnhanceNet forward path (BASELINE.json metric:
SR images/sec, 64x64 -> 256x256, bf16 tensor-core math).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch 64]
    python bench.py --workload train [--gpus N] ...      (BASELINE config 5, not the headline: see run_train)

One "step" = one forward of a batch of 64 synthetic 64x64 images (BASELINE.json config 2) per GPU.
N > 1 (torchrun, one rank per GPU, NCCL): every rank runs the same per-GPU batch on its own shard -
weak scaling, no data-path collective; the step time is the max over ranks.

Prints ONE JSON line (rank 0):
  value     images/s with inputs resident in HBM (inputs rotate through a pool larger than L2)
  e2e       images/s through the public API with pinned HOST buffers: H2D of the LR batch, forward,
            D2H of the SR batch, every step inside the timed region
  roofline  the dominant kernel (body2_umma_kernel: the 127 64->64 3x3 convolutions of the body in one
            persistent launch): algorithmic FLOPs per launch / its average launch time (CUDA events on
            the launch stream), against the measured sustained bf16 peak
  cpu_baseline  the fp32 CPU oracle (a port of the reference's forward) on this box's host cores
--impl reference: times that CPU path alone (the reference is pure Python/PyTorch and is not installed
on the GPU box; oracle/fen_oracle.py restates its forward with the same ATen ops).
"""
from __future__ import annotations

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

FLOP_PER_IMAGE = 44.633e9          # 22 316 703 744 MAC x 2, forward, 64x64 -> 256x256
CONV64_FLOP_PER_IMAGE = 2.0 * 4096 * 64 * 64 * 9   # one 64->64 3x3 conv on a 64x64 map
MODEL_CFG = dict(num_groups=6, blocks_per_group=10)
WORKLOAD = "FaceEnhanceNet 6x10x64 bf16 inference, batch 64/GPU, synthetic 64x64 -> 256x256 (BASELINE config 2)"


def ncu_traffic(batch: int):
    """DRAM bytes per launch of the body kernel from the committed `ncu --set full` capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "body_kernel_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["dram_bytes_per_launch"]) if int(d.get("batch", -1)) == batch else None
    except (OSError, ValueError, KeyError):
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_sustained": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                "bf16_burst": float(d.get("bf16_tflops", 1590.0)), "hbm": float(d.get("hbm_gbs", 6650.0)),
                "source": "MEASURED_PEAKS.json"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_forward_rate(batch: int, budget_s: float, min_iters: int = 2):
    """images/s of the fp32 CPU oracle (port of the reference forward) with all host threads."""
    import torch
    from oracle import fen_oracle, weights
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = weights.make_state_dict(0, "T1", **MODEL_CFG)
    x = torch.rand(batch, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        fen_oracle.fen_forward(sd, x)  # warm-up
        t0, n = time.perf_counter(), 0
        while n < min_iters or (time.perf_counter() - t0 < budget_s and n < 100):
            fen_oracle.fen_forward(sd, x)
            n += 1
        dt = time.perf_counter() - t0
    return batch * n / dt, threads, n, dt


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU forward (oracle port) on the host cores, rank 0 only."""
    if rank != 0:
        return
    import torch
    from oracle import fen_oracle, weights
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sample_batch = 8
    sd = weights.make_state_dict(0, "T1", **MODEL_CFG)
    x = torch.rand(sample_batch, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        for _ in range(max(1, args.warmup)):
            fen_oracle.fen_forward(sd, x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fen_oracle.fen_forward(sd, x)
        dt = time.perf_counter() - t0
    value = sample_batch * args.steps / dt
    sample = f"{sample_batch} images per step (bounded sample of the batch-64 workload), fp32, torch {torch.__version__}"
    line = {
        "impl": "reference", "metric": "SR images/sec (64->256)", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


TRAIN_WORKLOAD = ("Stage-1 L1 training step (float LR generation, forward, L1, backward, NCCL gradient all-reduce, "
                  "clip 0.5 + AdamW), FaceEnhanceNet 6x10x64, batch 32/GPU, random-init T1 weights (BASELINE config 5)")


def run_train(args, rank: int, local_rank: int, world: int):
    """--workload train: BASELINE config 5.  One step = training.Stage1Step.step on a batch of 32 synthetic
    256x256 HR images per GPU; data parallel under torchrun (one NCCL all-reduce of the flat gradient per step)."""
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    entry.build()
    import fsr_b200
    from fsr_b200 import _lib, data, sharding, training
    from oracle import fen_oracle, weights
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    warmup, B, lib = max(3, args.warmup), (32 if args.batch == 64 else args.batch), _lib.load()
    model = fsr_b200.FaceEnhanceNet(**MODEL_CFG)
    model.load_state_dict(weights.make_state_dict(0, "T1", **MODEL_CFG), strict=True)
    model = model.to(dev).train()
    stepper = fsr_b200.Stage1Step(model)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    pool = [torch.rand(B, 3, 256, 256, device=dev, generator=gen) for _ in range(8)]   # 8 x 25 MB > 126 MB L2
    # kernels per step, counted once by hand through the same calls Stage1Step.step makes
    lr_img, _ = data.lr_from_hr_float(pool[0]); n_launch = lib.fen_last_launch_count()
    sr, ws = model._forward_train(lr_img); n_launch += lib.fen_last_launch_count()
    _, dsr = training.l1_loss(sr, pool[0]); n_launch += 2
    model._backward(lr_img, dsr, ws); n_launch += lib.fen_last_launch_count() + 3
    for i in range(warmup):
        stepper.step(pool[i % 8])
    torch.cuda.synchronize(); sharding.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss, _ = stepper.step(pool[(warmup + i) % 8])
    e1.record(); torch.cuda.synchronize(); sharding.barrier()
    ms_total = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms_total * 1e-3)
    # e2e: pinned host HR batch -> H2D -> step -> loss read back on the host, every step
    host = [torch.rand(B, 3, 256, 256).pin_memory() for _ in range(2)]
    for i in range(warmup):
        stepper.step(host[i & 1].to(dev, non_blocking=True))[0].item()
    torch.cuda.synchronize(); sharding.barrier()
    e0.record()
    for i in range(args.steps):
        loss_host = stepper.step(host[i & 1].to(dev, non_blocking=True))[0].item()
    e1.record(); torch.cuda.synchronize(); sharding.barrier()
    ms_e2e = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    peaks = measured_peaks()
    step_tflops = value / world * 3 * FLOP_PER_IMAGE / 1e12
    line = {
        "metric": "Stage-1 training images/sec (L1, 64->256, bf16 activations, fp32 master weights)", "value": value,
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": TRAIN_WORKLOAD, "global_batch": world * B,
                   "parallelism": f"data parallel x{world}, one NCCL all-reduce of the 20.5 MB fp32 gradient per step",
                   "l2": "HR batches rotate through a 201 MB pool (> 126 MB L2); the step keeps 3.8 GB of activations"},
        "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "images/s",
                "h2d_bytes_per_step": B * 3 * 256 * 256 * 4, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                "last_loss": loss_host},
        "gpu_launches": n_launch * args.steps,
        "roofline": {"bound": "tensor", "kernel": "whole step (forward + backward = 3 x 44.633 GFLOP per image)",
                     "achieved": step_tflops, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": step_tflops / peaks["bf16_sustained"], "traffic": None,
                     "peak_source": peaks["source"] + " bf16_tflops_sustained"},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        sd = weights.make_state_dict(0, "T1", **MODEL_CFG)
        x = torch.rand(2, 3, 64, 64)
        dout = torch.full((2, 3, 256, 256), 1.0 / (2 * 3 * 256 * 256))
        fen_oracle.fen_backward(sd, x, dout)
        t0, n = time.perf_counter(), 0
        while n < 2 or (time.perf_counter() - t0 < 10.0 and n < 50):
            fen_oracle.fen_backward(sd, x, dout); n += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 2 * n / dt, "unit": "images/s", "cores": threads, "kind": "port",
                                "sample": f"{n} fp32 forward + backward passes of batch 2 in {dt:.1f} s (autograd oracle)"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="infer", choices=["infer", "train"],
                    help="infer = the headline (BASELINE config 2); train = the Stage-1 step (config 5)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload == "train":
        run_train(args, rank, local_rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    entry.build()
    import fsr_b200
    from fsr_b200 import _lib, sharding
    from oracle import weights

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(3, args.warmup)
    B = args.batch
    lib = _lib.load()

    model = fsr_b200.FaceEnhanceNet(**MODEL_CFG)
    model.load_state_dict(weights.make_state_dict(0, "T1", **MODEL_CFG), strict=True)
    model = model.to(dev).eval()

    # input pool larger than L2 (126 MB): 48 x 3.1 MB batches, a different one every step
    pool_n = 48
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    pool = [torch.rand(B, 3, 64, 64, device=dev, generator=gen) for _ in range(pool_n)]

    def step(i):
        with torch.no_grad():
            return model(pool[i % pool_n])

    for i in range(warmup):
        step(i)
    launches_per_step = lib.fen_last_launch_count()
    torch.cuda.synchronize()
    sharding.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        step(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    sharding.barrier()
    ms_total = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- e2e: pinned host LR batch -> H2D -> forward -> D2H of the SR batch, every step, through the
    # public API (model(x)).  The D2H of step i runs on a copy stream and overlaps the forward of step
    # i+1 (two pinned output buffers); every byte of every step is still moved inside the timed region.
    host_in = [torch.rand(B, 3, 64, 64).pin_memory() for _ in range(4)]
    host_out = [torch.empty(B, 3, 256, 256).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    done = [torch.cuda.Event() for _ in range(2)]

    def e2e_step(i):
        with torch.no_grad():
            x = host_in[i % 4].to(dev, non_blocking=True)
            y = model(x)
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ready)
                done[i & 1].synchronize() if i >= 2 else None   # pinned buffer i&1 is free again
                host_out[i & 1].copy_(y, non_blocking=True)
                y.record_stream(copy_stream)
                done[i & 1].record(copy_stream)

    for i in range(warmup):
        e2e_step(i)
    torch.cuda.synchronize()
    sharding.barrier()
    e0.record()
    for i in range(args.steps):
        e2e_step(i)
    torch.cuda.current_stream().wait_stream(copy_stream)
    e1.record()
    torch.cuda.synchronize()
    sharding.barrier()
    ms_e2e = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel: body2_umma_kernel (all 127 64->64 3x3 convs of the body in one
    # persistent launch, 86 % of the FLOPs).  Its launches are timed with CUDA events recorded on the
    # launch stream by the library itself (fen_profile_body) during extra forwards of the same workload.
    n_body_convs = MODEL_CFG["num_groups"] * (2 * MODEL_CFG["blocks_per_group"] + 1) + 1
    lib.fen_profile_body(1)
    body_ms = []
    for i in range(max(5, min(args.steps, 20))):
        step(i)
        body_ms.append(lib.fen_last_body_ms())
    lib.fen_profile_body(0)
    body_ms = [m for m in body_ms if m > 0]
    peaks = measured_peaks()
    if body_ms:
        k_ms = sum(body_ms) / len(body_ms)
        k_name = "body2_umma_kernel (127 x 64->64 3x3 conv + SE + residuals in one persistent launch, batch %d)" % B
        k_flop = CONV64_FLOP_PER_IMAGE * B * n_body_convs
    else:  # configurations the persistent kernel does not cover fall back to per-layer launches
        act = [torch.randn(B, 64, 64, 64, device=dev).mul_(0.3).to(torch.bfloat16) for _ in range(2)]
        w = torch.randn(64, 64, 3, 3, device=dev) * 0.05
        wp = torch.empty(9 * 64 * 64, dtype=torch.bfloat16, device=dev)
        bias = torch.zeros(64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.fen_pack_conv3x3(w.data_ptr(), 64, 64, wp.data_ptr(), st), "fen_pack_conv3x3")
        for i in range(46):
            if i == 6:
                torch.cuda.synchronize()
                e0.record()
            _lib.check(lib.fen_conv3x3_c64(act[i & 1].data_ptr(), wp.data_ptr(), bias.data_ptr(), None, None, None,
                                           act[(i + 1) & 1].data_ptr(), B, 64, 64, 5, st), "fen_conv3x3_c64")
        e1.record()
        torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / 40
        k_name = "conv3x3_umma_kernel<64> (64->64 3x3 conv, batch %d)" % B
        k_flop = CONV64_FLOP_PER_IMAGE * B
    conv_tflops = k_flop / (k_ms * 1e-3) / 1e12
    conv_ms = k_ms
    step_tflops = value / world * FLOP_PER_IMAGE / 1e12

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        rate, threads, n, dt = cpu_forward_rate(batch=1, budget_s=12.0)
        cpu = {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"{n} fp32 forwards of batch 1 in {dt:.1f} s (protocol of scripts/measure_inference_time.py:68-116)"}

    line = {
        "metric": "SR images/sec (64->256, bf16)", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": world * B, "parallelism": f"batch-sharded x{world}, no collective",
                   "l2": f"inputs rotate through a {pool_n * B * 3 * 64 * 64 * 4 / 1e6:.0f} MB pool (> 126 MB L2); "
                         "per-step activation working set ~1 GB"},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * 64 * 64 * 4,
                "d2h_bytes_per_step": B * 3 * 256 * 256 * 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "tensor", "kernel": k_name,
                     "achieved": conv_tflops, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": conv_tflops / peaks["bf16_sustained"], "traffic": ncu_traffic(B) if body_ms else None,
                     "peak_source": peaks["source"] + " bf16_tflops_sustained", "us_per_launch": conv_ms * 1e3,
                     "step_achieved": step_tflops, "step_frac": step_tflops / peaks["bf16_sustained"]},
        "clocks": clocks,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
