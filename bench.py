#!/usr/bin/env python
"""bench.py — denoising-syn training throughput (128x128 patches / s), the metric of BASELINE.json.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--precision bf16|tf32]
  torchrun --nproc-per-node N bench.py --gpus N ...          (one rank per GPU, NCCL)
  python bench.py --impl reference ...                        (the reference's CPU path, oracle port)

A "step" is one pass of the hot path over one batch of synthetic patches per GPU: forward
(SNet + RNet), fused ELBO, backward (dgrad + wgrad), gradient all-reduce (N > 1), per-sub-net
clip + Adam.  Prints ONE JSON line on rank 0 (see README / DESIGN.md §measurement).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

TRAIN_GFLOP_PER_PATCH = 245.15      # fwd + dgrad + wgrad, 2*MAC over conv layers
FWD_GFLOP_PER_PATCH = 81.735
N_FEAT, N_RES, DEP_S = [96, 192, 288], 3, 5
PATCH = 128
ALPHA0, EPS2 = 24.5, 1e-6           # 0.5 * var_window**2, configs/denoising_syn.json:37-38


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def synth_batch(b, rank, device, pinned=False):
    """U[0,1) clean patch, Gaussian-bump sigma map (SimulateTrain.generate_sigma_niid,
    datasets/DenoisingDatasets.py:190-203), noisy = gt + randn * sigma."""
    import torch
    g = torch.Generator().manual_seed(1000 + rank)
    im_gt = torch.rand(b, 3, PATCH, PATCH, generator=g)
    ii, jj = torch.meshgrid(torch.arange(PATCH, dtype=torch.float32), torch.arange(PATCH, dtype=torch.float32),
                            indexing="ij")
    ch = torch.rand(b, 1, 1, 1, generator=g) * PATCH
    cw = torch.rand(b, 1, 1, 1, generator=g) * PATCH
    s = 32 + torch.rand(b, 1, 1, 1, generator=g) * 64
    bump = torch.exp(-((ii - ch) ** 2 + (jj - cw) ** 2) / (2 * s ** 2))
    lo_hi = torch.sort(torch.rand(b, 2, generator=g) * 75 / 255, dim=1).values
    lo, hi = lo_hi[:, :1, None, None], lo_hi[:, 1:, None, None] + 5 / 255
    mn, mx = bump.amin(dim=(2, 3), keepdim=True), bump.amax(dim=(2, 3), keepdim=True)
    sig = lo + (bump - mn) / (mx - mn) * (hi - lo)
    im_noisy = im_gt + torch.randn(b, 3, PATCH, PATCH, generator=g) * sig
    sigma_gt = (sig ** 2).clamp_min(1e-10)
    ts = [im_noisy.contiguous(), im_gt.contiguous(), sigma_gt.contiguous()]
    if pinned:
        ts = [t.pin_memory() for t in ts]
    elif device is not None:
        ts = [t.to(device) for t in ts]
    return ts


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.samples:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------
# the reference's CPU path (oracle port), timed on the host cores
# --------------------------------------------------------------------------------------
def cpu_reference_steps(steps, warmup, b):
    import torch
    from oracle import virnet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.NetCfg(n_feat=tuple(N_FEAT), n_resblocks=N_RES, dep_S=DEP_S)
    torch.manual_seed(1234)
    sd = O.build_state_dict(cfg)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)
    pR = [v for k, v in params.items() if "rnet" in k.lower()]
    pS = [v for k, v in params.items() if "snet" in k.lower()]
    im_noisy, im_gt, sigma_gt = synth_batch(b, 0, None)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        mu, sigma = O.vir_denoise_forward(params, im_noisy, cfg)
        loss, *_ = O.elbo_denoising_simple(mu, sigma, im_noisy, im_gt, EPS2, ALPHA0, ALPHA0 * sigma_gt)
        loss.backward()
        O.clip_grad_norm_(pR, 1e3)
        O.clip_grad_norm_(pS, 1e2)
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return b / sec, sec, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    b = 2
    steps, warmup = min(args.steps, 10), min(args.warmup, 2)
    v, sec, cores = cpu_reference_steps(steps, warmup, b)
    line = {
        "impl": "reference", "metric": "denoising-syn 128x128 training patches/s (whole job)", "value": v,
        "unit": "patches/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "train_denoising_syn.py step (fwd+ELBO+bwd+clip+Adam), VIRAttResUNet "
                               "n_feat=[96,192,288] n_resblocks=3 dep_S=5, 128x128x3 patches", "batch_per_step": b},
        "cpu_baseline": {"value": v, "unit": "patches/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps of batch {b} (oracle/virnet_oracle.py = reference ops on torch CPU, "
                                   f"{cores} threads)"},
        "e2e": {"value": v, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
class _StdoutToStderr:
    """Native libraries (NCCL prints its version banner) write to file descriptor 1; the driver wants exactly ONE
    JSON line on stdout, so fd 1 points at stderr while the benchmark runs and is restored for the final print."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def run_ours(args):
    with _StdoutToStderr():
        line = _run_ours(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def _run_ours(args):
    import torch
    import torch.distributed as dist
    import virnet_b200
    from virnet_b200 import lib, ops
    from virnet_b200.trainer import DenoiseTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    b = args.batch
    torch.manual_seed(1234)                                            # train_denoising_syn.py:52-53
    net = virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=N_FEAT, dep_S=DEP_S, n_resblocks=N_RES,
                                    noise_cond=True, extra_mode="Input", noise_avg=False,
                                    precision=args.precision).to(dev)
    trainer = DenoiseTrainer(net, lr=1e-4, clip_grad_R=1e3, clip_grad_S=1e2, alpha0=ALPHA0, eps2=EPS2)
    host = synth_batch(b, rank, None, pinned=True)
    resident = [t.to(dev) for t in host]
    h2d = sum(t.numel() * 4 for t in host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    loss_host = [torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    e2e_i = [0]
    e2e_last = [float('nan')]

    def step_resident():
        trainer.step(*resident)

    def step_e2e():
        # every step copies its inputs from pinned host memory (H2D inside the timed region); like a DataLoader with
        # pin_memory, the copy of the NEXT step's batch is started while this step computes
        losses = trainer.step(*host)
        trainer.prefetch(*host)
        # D2H read of THIS step's result, issued in stream order; the host consumes it one step later (asynchronous
        # logging), so the GPU never waits for Python to enqueue the next step.  All reads land before the final sync.
        i = e2e_i[0]
        loss_host[i & 1].copy_(losses, non_blocking=True)
        loss_ev[i & 1].record()
        if i > 0:
            loss_ev[(i - 1) & 1].synchronize()
            e2e_last[0] = float(loss_host[(i - 1) & 1][0])
        e2e_i[0] = i + 1

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)
    l0 = lib.launch_count()
    t0 = time.time()
    ms = timed(step_resident, args.steps)
    t1 = time.time()
    launches = lib.launch_count() - l0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    torch.cuda.synchronize()
    final_loss = float(loss_host[(e2e_i[0] - 1) & 1][0])

    # ---- roofline of the dominant kernel (conv_igemm: fprop + dgrad launches), measured live ----
    # (per-launch CUDA events need the launches serialised on one stream: the weight-gradient side stream of
    # the product path is switched off for these two extra steps only; `value` / `e2e` above ran with it on)
    side = trainer.engine.wgrad_side_stream
    trainer.engine.wgrad_side_stream = False
    prof = ops.start_profile()
    for _ in range(2):
        step_resident()
    torch.cuda.synchronize()
    recs = ops.stop_profile()
    trainer.engine.wgrad_side_stream = side
    fam = {}
    for r in recs:
        f = fam.setdefault(r["family"], {"ms": 0.0, "flops": 0.0, "n": 0})
        f["ms"] += r["ms"]; f["flops"] += r["flops"]; f["n"] += 1
    peaks, peak_src = load_peaks()
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) * (0.5 if args.precision == "tf32" else 1.0)
    conv = fam.get("conv_igemm", {"ms": 1.0, "flops": 0.0, "n": 1})
    achieved = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        traffic = json.loads(tp.read_text()).get("conv_igemm_dram_bytes_per_launch")
    step_ms_prof = sum(f["ms"] for f in fam.values()) / 2.0

    value = world * b * args.steps / (ms * 1e-3)
    e2e = world * b * args.steps / (ms_e2e * 1e-3)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores = cpu_reference_steps(2, 1, 2)
        cpu = {"value": v, "unit": "patches/s", "cores": cores, "kind": "port",
               "sample": f"2 timed steps of batch 2 after 1 warm-up (oracle port of the reference, torch CPU, {cores} threads)"}
    if rank == 0:
        line = {
            "metric": "denoising-syn 128x128 training patches/s (whole job)", "value": value, "unit": "patches/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "train_denoising_syn.py step (fwd+ELBO+bwd+allreduce+clip+Adam), VIRAttResUNet "
                                   "n_feat=[96,192,288] n_resblocks=3 dep_S=5, 128x128x3 patches (configs[2])",
                       "batch_per_gpu": b, "global_batch": b * world, "parallelism": f"dp{world}",
                       "l2": "working set per step (activations >1 GB) exceeds the 126 MB L2; no flush needed",
                       "train_gflop_per_patch": TRAIN_GFLOP_PER_PATCH, "final_loss": final_loss},
            "e2e": {"value": e2e, "unit": "patches/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "vk_conv_igemm launches (conv_v2_kernel + conv_igemm_kernel; fprop+dgrad), serialised",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": f"{peak_src} bf16_tflops_sustained" + (" x0.5 (tf32)" if args.precision == "tf32" else ""),
                         "traffic": traffic, "launches_per_step": conv["n"] // 2,
                         "share_of_step": (conv["ms"] / 2.0) / step_ms_prof if step_ms_prof else None,
                         "families_ms_per_step": {k: round(v["ms"] / 2.0, 4) for k, v in fam.items()},
                         "whole_step_tflops": value * TRAIN_GFLOP_PER_PATCH / 1e3 / world},
            "cpu_baseline": cpu,
        }
    else:
        line = None
    if world > 1:
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=32,
                    help="patches per GPU per step (weak scaling; the reference trains with a GLOBAL batch of 16)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
