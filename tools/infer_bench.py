#!/usr/bin/env python
"""Secondary measurements (not the headline metric): inference throughput of configs[1] (denoising-syn 256x256,
batch 32) and configs[3]-shaped x4 super-resolution, through the drop-in nn.Module API, CUDA-event timed."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import virnet_b200  # noqa: E402


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


out = []
for prec in ("bf16", "tf32"):
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=[96, 192, 288], dep_S=5, n_resblocks=3,
                                    noise_cond=True, extra_mode="Input", noise_avg=False, precision=prec).cuda().eval()
    x = torch.rand(32, 3, 256, 256, device="cuda")
    with torch.no_grad():
        ms = timed(lambda: net(x))
    out.append(dict(workload="denoising-syn inference 256x256 batch 32 (configs[1])", precision=prec, ms_per_batch=ms,
                    images_per_s=32 / ms * 1e3, fwd_tflops=32 * 326.94 / ms))
    torch.manual_seed(1234)
    sr = virnet_b200.VIRAttResUNetSR(im_chn=3, sigma_chn=1, dep_S=5, dep_K=8, n_feat=[96, 160, 224], n_resblocks=2,
                                     extra_mode="Both", noise_avg=True, noise_cond=True, kernel_cond=True,
                                     precision=prec).cuda().eval()
    lr = torch.rand(16, 3, 64, 64, device="cuda")
    with torch.no_grad():
        ms = timed(lambda: sr(lr, 4))
    out.append(dict(workload="sisr x4 inference LR 64x64 -> 256x256 batch 16", precision=prec, ms_per_batch=ms,
                    images_per_s=16 / ms * 1e3, fwd_tflops=16 * 180.16 / ms))
for o in out:
    print(json.dumps(o))
