#!/usr/bin/env python
"""Eager vs CUDA-graph training step at small per-GPU batches (the reference's own global batch of 16 is 2 patches per
GPU on 8 GPUs): patches/s and ms/step on one GPU.  usage: graph_bench.py [batches...]"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import bench  # noqa: E402
import virnet_b200  # noqa: E402
from virnet_b200.trainer import DenoiseTrainer  # noqa: E402

dev = torch.device("cuda", 0)
for b in [int(a) for a in sys.argv[1:]] or [2, 4, 16, 32]:
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=bench.N_FEAT, dep_S=bench.DEP_S, n_resblocks=bench.N_RES,
                                    noise_cond=True, extra_mode="Input", noise_avg=False, precision="bf16").to(dev)
    tr = DenoiseTrainer(net)
    batch = bench.synth_batch(b, 0, dev)
    res = {}
    for name, fn in (("eager", tr.step), ("graph", tr.step_graph)):
        for _ in range(5):
            fn(*batch)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(30):
            fn(*batch)
        e.record()
        torch.cuda.synchronize()
        res[name] = s.elapsed_time(e) / 30
    print(json.dumps(dict(batch=b, eager_ms=round(res["eager"], 3), graph_ms=round(res["graph"], 3),
                          eager_patches_s=round(b / res["eager"] * 1e3, 1), graph_patches_s=round(b / res["graph"] * 1e3, 1))),
          flush=True)
