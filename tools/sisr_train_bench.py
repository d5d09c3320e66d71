#!/usr/bin/env python
"""Times the SISR training step (BASELINE.json configs[4]: train_SISR x4, 64x64 -> 256x256 patches, the
shipped sisr_x4.json network) on one GPU and prints a per-family breakdown of one step.  Tuning aid.
usage: sisr_train_bench.py [batch] [precision] [lr_size] [--graph] [--det]   (--det: SISRTrainer(deterministic=True))"""
import json
import sys
from collections import defaultdict
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import virnet_b200  # noqa: E402
from virnet_b200 import ops  # noqa: E402
from virnet_b200.trainer import SISRTrainer  # noqa: E402

_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(_pos[0]) if len(_pos) > 0 else 16
prec = _pos[1] if len(_pos) > 1 else "bf16"
lr_sz = int(_pos[2]) if len(_pos) > 2 else 64
sf = 4
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
net = virnet_b200.VIRAttResUNetSR(im_chn=3, sigma_chn=1, kernel_chn=3, n_feat=[96, 160, 224], dep_S=5, dep_K=8,
                                  noise_cond=True, kernel_cond=True, n_resblocks=2, extra_mode="Both", noise_avg=True,
                                  precision=prec).to(dev).train()
g = torch.Generator(device=dev).manual_seed(0)
im_hr = torch.rand(B, 3, lr_sz * sf, lr_sz * sf, device=dev, generator=g)
im_lr = torch.nn.functional.avg_pool2d(im_hr, sf) + 0.01 * torch.randn(B, 3, lr_sz, lr_sz, device=dev, generator=g)
kinfo_gt = torch.stack([0.5 + 3 * torch.rand(B, device=dev, generator=g), 0.5 + 3 * torch.rand(B, device=dev, generator=g),
                        torch.rand(B, device=dev, generator=g) - 0.5], dim=1)
nlevel = torch.full((B, 1, 1, 1), (2.55 / 255) ** 2, device=dev)
DET = "--det" in sys.argv
tr = SISRTrainer(net, sf, deterministic=DET)
for _ in range(3):
    tr.step(im_hr, im_lr, kinfo_gt, nlevel)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 10
s.record()
for _ in range(iters):
    terms = tr.step(im_hr, im_lr, kinfo_gt, nlevel)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / iters
print(json.dumps(dict(workload=f"train_SISR x{sf} {lr_sz}->{lr_sz * sf} b={B} {prec}" + (" deterministic" if DET else ""),
                      ms_per_step=round(ms, 3),
                      patches_per_s=round(B / ms * 1e3, 1), loss=terms[0].item())))
if "--graph" in sys.argv:
    for _ in range(3):
        tr.step_graph(im_hr, im_lr, kinfo_gt, nlevel)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        terms = tr.step_graph(im_hr, im_lr, kinfo_gt, nlevel)
    e.record()
    torch.cuda.synchronize()
    msg = s.elapsed_time(e) / iters
    print(json.dumps(dict(workload=f"train_SISR x{sf} {lr_sz}->{lr_sz * sf} b={B} {prec} CUDA graph", ms_per_step=round(msg, 3),
                          patches_per_s=round(B / msg * 1e3, 1), loss=terms[0].item())))
    sys.exit(0)
tr.engine.wgrad_side_stream = False
tr.step(im_hr, im_lr, kinfo_gt, nlevel)
ops.start_profile()
tr.step(im_hr, im_lr, kinfo_gt, nlevel)
recs = ops.stop_profile()
agg = defaultdict(lambda: [0.0, 0.0, 0])
for r in recs:
    key = r["family"] if r["flops"] == 0 else f'{r["family"]} {r["flops"] / 1e9:7.2f}GF'
    agg[key][0] += r["ms"]
    agg[key][1] += r["flops"]
    agg[key][2] += 1
tot = sum(v[0] for v in agg.values())
print(f"serialised launches total {tot:.3f} ms")
for fam, (m, fl, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    tf = fl / (m * 1e-3) / 1e12 if m > 0 else 0
    print(f"{fam:28s} n={n:4d} ms={m:8.3f} ({100 * m / tot:4.1f}%) {tf:7.0f} TF/s")
