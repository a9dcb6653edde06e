#!/usr/bin/env python
"""Headline benchmark: UNet 2-D 256x256x3 fwd+bwd+optimizer training step, bf16, batch 64 per GPU.

  python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels)
  python bench.py --impl reference --steps K --warmup W    # the reference path's CPU stand-in

Metric (BASELINE.json): UNet fwd+bwd slices/s @256^2; conv TFLOP/s vs tensor-core peak.
 * `value`   = slices/s over all N GPUs, inputs resident in HBM, K steps timed with CUDA events on the
               compute stream, bracketed by barrier + stream sync, MAX over ranks.
 * `e2e`     = the same step through the host-fed call (pinned host batch -> H2D -> step -> D2H loss).
 * `roofline`= all tcgen05 conv launches of the timed steps: algorithmic FLOPs / summed event durations,
               against the measured bf16 peak in MEASURED_PEAKS.json (sustained figure: timed inside a long step).
 * `cpu_baseline` / `--impl reference`: the reference's own implementation is TensorFlow 1.13, which cannot
               be installed in this image (Python 3.12, no network). The stand-in is the
               numpy/BLAS oracle port of the same training step (oracle/unet_ref.py) on the box's host cores.
Weak scaling: per-GPU batch fixed at 64; gradients are all-reduced with NCCL every step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HW = 256
BATCH_PER_GPU = 64
FLOP_PER_SLICE = 288.828e9  # BASELINE.md section 3, UNet 2-D 256x256x3, 3 classes, fwd+bwd


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------
def cpu_port_throughput(target_seconds: float = 15.0, batch: int = 1, steps: int | None = None, warmup: int = 0):
    """Times oracle.unet_ref.train_step (fp32, numpy + BLAS, all host threads) on `batch` slices of 256^2."""
    import numpy as np
    from boxsegliver_b200 import synthetic
    from oracle import unet_ref as R

    cfg = R.UNetCfg(height=HW, width=HW, channel=3, loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4),
                    weight_decay_rate=1e-6)
    params = R.init_params(cfg, seed=0)
    slots = {}
    images, labels = synthetic.make_batch(batch, HW, HW, 3)
    for i in range(warmup):
        R.train_step(params, slots, i + 1, images, labels, cfg, 1e-3)
    times = []
    t_all = time.perf_counter()
    i = 0
    while True:
        t0 = time.perf_counter()
        R.train_step(params, slots, warmup + i + 1, images, labels, cfg, 1e-3)
        times.append(time.perf_counter() - t0)
        i += 1
        if steps is not None:
            if i >= steps:
                break
        elif time.perf_counter() - t_all > target_seconds or i >= 8:
            break
    sec = sum(times) / len(times)
    return {"slices_per_s": batch / sec, "sec_per_step": sec, "steps": len(times), "batch": batch}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = cpu_port_throughput(batch=1, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    cores = host_cores()
    sample = (f"{r['steps']} training steps of batch {r['batch']} (UNet 2-D 256x256x3, fp32, BN, weighted xent, "
              f"L2, Adam) with the numpy/BLAS oracle port; TensorFlow 1.13 is not installable here")
    line = {
        "impl": "reference", "metric": "unet2d_256_train_slices_per_s", "value": r["slices_per_s"], "unit": "slices/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": min(args.warmup, 1),
        "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "UNet 2D 256x256x3 training step (fwd+bwd+Adam), bounded sample: batch 1 per step",
                   "per_gpu_batch": BATCH_PER_GPU, "classes": 3, "normalizer": "batch_norm"},
        "cpu_baseline": {"value": r["slices_per_s"], "unit": "slices/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": r["slices_per_s"], "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="bsl_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        try:
            with open(self.path) as f:
                for ln in f:
                    parts = [p.strip() for p in ln.split(",")]
                    if len(parts) < 9:
                        continue
                    try:
                        a, b, c = float(parts[1]), float(parts[2]), float(parts[3])
                    except ValueError:
                        continue
                    sm.append(a)
                    mx.append(b)
                    pw.append(c)
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                       parts[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            # "under load" = samples drawing at least 60 % of the highest power seen in the window
            lim = 0.6 * max(pw)
            load = [a for a, c in zip(sm, pw) if c >= lim] or sm
            out.update(sm_mhz=statistics.median(load), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       samples_under_load=len(load), power_w_max=max(pw))
        return out


def run_b200(args):
    import ctypes as C

    import numpy as np

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dist = None
    if world > 1:
        import torch.distributed as dist  # control plane only: id broadcast, barriers, max over ranks
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)

    from boxsegliver_b200 import synthetic
    from boxsegliver_b200.device import Context
    from boxsegliver_b200.engine import EngineConfig, UNetEngine

    ctx = Context(local)
    cfg = EngineConfig(batch=BATCH_PER_GPU, height=HW, width=HW, channel=3, classes=("Background", "Liver", "Tumor"),
                       normalizer="batch_norm", weight_decay_rate=1e-6, loss_type="xentropy",
                       loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4), optimizer="adam", world=world)
    eng = UNetEngine(ctx, cfg)
    eng.init_weights(seed=0)  # identical on every rank (mirrored variables)
    if world > 1:
        uid = (C.c_char * 128)()
        if rank == 0:
            ctx.call("bsl_comm_unique_id", uid)
        box = [bytes(uid)]
        dist.broadcast_object_list(box, src=0)
        eng.attach_comm(rank, world, box[0])
    images, labels = synthetic.make_batch(BATCH_PER_GPU, HW, HW, 3, seed=1357 + rank)
    eng.set_inputs(images, labels)
    lr = 1e-3

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    def launches():
        n = C.c_ulonglong()
        ctx.call("bsl_launch_count", C.byref(n))
        return n.value

    # ---- device-resident arm
    sampler = ClockSampler(local) if rank == 0 else None   # started before warm-up: nvidia-smi takes ~1 s to spin up
    for _ in range(max(args.warmup, 3)):
        eng.train_step(lr)
    barrier()
    l0 = launches()
    e0, e1 = ctx.new_event(), ctx.new_event()
    ctx.record(e0)
    for _ in range(args.steps):
        eng.train_step(lr)
    ctx.record(e1)
    ms_total = ctx.elapsed_ms(e0, e1)
    barrier()
    n_launch = launches() - l0
    # Roofline pass: the same K steps again with every tensor-core launch bracketed by CUDA events on its stream and
    # the filter-gradient side stream folded back into the compute stream, so that each kernel is timed ALONE (with
    # the overlap on, a bracket also contains the time the kernel spends sharing SMs with the normalisation backward).
    eng.enable_conv_timing(True)
    overlap_was, eng._overlap_wgrad = eng._overlap_wgrad, False
    for _ in range(args.steps):
        eng.train_step(lr)
    conv = eng.conv_timing_report()
    eng.enable_conv_timing(False)
    eng._overlap_wgrad = overlap_was
    clocks = sampler.stop() if sampler else None
    ctx.check_device()

    # ---- end-to-end arm: host batches in (pinned memory -> H2D every step), loss out (D2H every step). The copy of
    # batch i + 1 is submitted before step i is launched, on a copy stream, and overlaps its compute.
    for j in (0, 1):
        si, sl = eng.staging_slot(j)
        si[...] = images
        sl[...] = labels
    eng.submit_staged(0)
    for i in range(2):
        eng.submit_staged((i + 1) % 2)
        eng.train_step_prefetched(lr)
    barrier()
    t0 = time.perf_counter()
    ctx.record(e0)
    loss = 0.0
    for i in range(args.steps):
        eng.submit_staged((i + 1) % 2)          # next batch's H2D (K copies for K timed steps)
        loss = eng.train_step_prefetched(lr)    # waits for its own batch's copy, runs the step, reads the loss back
    ctx.record(e1)
    e2e_ms_total = ctx.elapsed_ms(e0, e1)
    e2e_wall = (time.perf_counter() - t0) * 1e3
    e2e_ms_total = max(e2e_ms_total, e2e_wall)  # the host waits on the loss every step: wall clock is the honest one
    ctx.sync(eng.copy_stream)
    barrier()
    ctx.check_device()

    if dist is not None:
        import torch
        t = torch.tensor([ms_total, e2e_ms_total], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms_total = float(t[0]), float(t[1])

    if rank == 0:
        peaks, peak_src = measured_peaks()
        ms = ms_total / args.steps
        value = world * BATCH_PER_GPU * args.steps / (ms_total / 1e3)
        e2e_value = world * BATCH_PER_GPU * args.steps / (e2e_ms_total / 1e3)
        achieved = conv["flops"] / (conv["ms"] / 1e3) / 1e12 if conv["ms"] > 0 else 0.0
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
        fl = eng.step_flops()
        traffic, traffic_note = None, None
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):   # dram__bytes_read+write of one ncu --set full capture (never measured under this run)
            with open(tp) as f:
                tj = json.load(f)
            traffic = tj.get("dram_bytes_per_launch")
            traffic_note = {"kernel": tj.get("dominant"), "algorithmic_bytes_per_launch": tj.get("algorithmic_bytes_per_launch"),
                            "source": tj.get("source")}
        per_kind = {k: {"launches_per_step": v[0] // args.steps, "ms_per_step": v[1] / args.steps,
                        "tflops": v[2] / (v[1] / 1e3) / 1e12 if v[1] > 0 else 0.0} for k, v in conv["per_kind"].items()}
        line = {
            "metric": "unet2d_256_train_slices_per_s", "value": value, "unit": "slices/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "UNet 2D 256x256x3 bf16 training (fwd+loss+bwd+allreduce+Adam), batch 64 per GPU "
                                   "(BASELINE.json configs[1])",
                       "per_gpu_batch": BATCH_PER_GPU, "global_batch": BATCH_PER_GPU * world, "classes": 3,
                       "normalizer": "batch_norm", "loss": "weighted xent 0.2/0.4/4.4 + L2 1e-6", "optimizer": "adam",
                       "parallelism": f"dp{world}", "wgrad_overlap": bool(eng._overlap_wgrad),
                       "l2_cache": "inputs larger than L2: ~20 GB of activations/gradients touched per step"},
            "model_tflops_per_s": FLOP_PER_SLICE * value / 1e12,
            "step_tflop_algorithmic": fl["total"] / 1e12,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "traffic_note": traffic_note, "per_kind": per_kind,
                         "kernel": "bsl::conv_halo_kernel + bsl::wgrad_halo_kernel + bsl::igemm_kernel (every tcgen05 conv / convT fprop, dgrad, wgrad launch of the timed steps)",
                         "launches_timed": conv["launches"], "ms_per_step_in_kernel": conv["ms"] / args.steps,
                         "peak_source": f"{peak_src} MEASURED_PEAKS.json bf16_tflops_sustained",
                         "timing": "CUDA events around every tensor-core launch over K extra steps run with the "
                                   "filter-gradient overlap off (kernels timed alone); `value` is measured with it on",
                         "whole_step_frac_of_peak": FLOP_PER_SLICE * value / 1e12 / world / peak},
            "e2e": {"value": e2e_value, "unit": "slices/s", "h2d_bytes_per_step": eng.h2d_bytes_per_step(),
                    "d2h_bytes_per_step": 12, "last_loss": loss},
            "gpu_launches": int(n_launch),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_port_throughput(target_seconds=12.0, batch=1)
            line["cpu_baseline"] = {
                "value": r["slices_per_s"], "unit": "slices/s", "cores": host_cores(), "kind": "port",
                "sample": f"{r['steps']} training step(s) of batch 1 at 256x256x3 with the numpy/BLAS oracle port "
                          f"({r['sec_per_step']:.2f} s/step); the reference's TF 1.13 cannot be installed here"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
