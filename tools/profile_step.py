#!/usr/bin/env python
"""Run a few denoising-syn training steps (bench.py's workload, nothing else) — the command to wrap
in ncu:  ncu --metrics gpu__time_duration.sum --clock-control none -s <launches of warm-up> -c <N> ...
         python tools/profile_step.py --steps 3 --batch 16 --precision bf16
Prints the number of library launches per step so -s / -c can be chosen."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import bench  # noqa: E402
import virnet_b200  # noqa: E402
from virnet_b200 import lib  # noqa: E402
from virnet_b200.trainer import DenoiseTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--fwd-only", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
net = virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=bench.N_FEAT, dep_S=bench.DEP_S, n_resblocks=bench.N_RES,
                                noise_cond=True, extra_mode="Input", noise_avg=False, precision=args.precision).to(dev)
batch = bench.synth_batch(args.batch, 0, dev)
if args.fwd_only:
    with torch.no_grad():
        for i in range(args.steps):
            l0 = lib.launch_count()
            net(batch[0])
            torch.cuda.synchronize()
            print(f"fwd {i}: {lib.launch_count() - l0} launches", flush=True)
else:
    tr = DenoiseTrainer(net, lr=1e-4, clip_grad_R=1e3, clip_grad_S=1e2, alpha0=bench.ALPHA0, eps2=bench.EPS2)
    for i in range(args.steps):
        l0 = lib.launch_count()
        losses = tr.step(*batch)
        torch.cuda.synchronize()
        print(f"step {i}: {lib.launch_count() - l0} launches, loss {losses[0].item():.4f}", flush=True)
