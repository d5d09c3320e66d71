#!/usr/bin/env python
"""Per-launch breakdown of one training step (CUDA events around every library call)."""
import sys
from collections import defaultdict
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import bench  # noqa: E402
import virnet_b200  # noqa: E402
from virnet_b200 import ops  # noqa: E402
from virnet_b200.trainer import DenoiseTrainer  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(1234)
net = virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=bench.N_FEAT, dep_S=bench.DEP_S, n_resblocks=bench.N_RES,
                                noise_cond=True, extra_mode="Input", noise_avg=False, precision="bf16").to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
batch = bench.synth_batch(B, 0, dev)
tr = DenoiseTrainer(net)
tr.engine.wgrad_side_stream = False   # serialise launches so per-launch events are meaningful
for _ in range(3):
    tr.step(*batch)
ops.start_profile()
tr.step(*batch)
recs = ops.stop_profile()
agg = defaultdict(lambda: [0.0, 0.0, 0])
for r in recs:
    key = (r["family"], round(r["flops"] / 1e9, 2))
    agg[key][0] += r["ms"]
    agg[key][1] += r["flops"]
    agg[key][2] += 1
tot = sum(v[0] for v in agg.values())
print(f"total {tot:.3f} ms")
for (fam, gf), (ms, fl, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    tf = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0
    print(f"{fam:14s} GFLOP/launch={gf:9.2f} n={n:3d} ms={ms:7.3f} ({100 * ms / tot:4.1f}%) {tf:7.0f} TF/s")
