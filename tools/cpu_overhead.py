#!/usr/bin/env python
"""Host-side enqueue time of one training step vs its GPU time (is the step launch-bound?)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import bench  # noqa: E402
import virnet_b200  # noqa: E402
from virnet_b200.trainer import DenoiseTrainer  # noqa: E402

dev = torch.device("cuda", 0)
for b in (16, 32):
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=bench.N_FEAT, dep_S=bench.DEP_S, n_resblocks=bench.N_RES,
                                    noise_cond=True, extra_mode="Input", noise_avg=False, precision="bf16").to(dev)
    batch = bench.synth_batch(b, 0, dev)
    tr = DenoiseTrainer(net)
    for _ in range(5):
        tr.step(*batch)
    torch.cuda.synchronize()
    n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        tr.step(*batch)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    # enqueue-only time: synchronise before every step so the queue is empty
    enq = 0.0
    for _ in range(n):
        torch.cuda.synchronize()
        a = time.perf_counter()
        tr.step(*batch)
        enq += time.perf_counter() - a
    torch.cuda.synchronize()
    print(f"b={b}: wall {1e3 * (t2 - t0) / n:.3f} ms/step, host enqueue (queue empty) {1e3 * enq / n:.3f} ms/step, "
          f"host loop without final sync {1e3 * (t1 - t0) / n:.3f} ms/step")
