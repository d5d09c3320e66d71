#!/usr/bin/env python
"""Achieved HBM bandwidth of the data-side / loss-side kernels added around the hot path, against their algorithmic
bytes: vk_synth_denoise, vk_noise_estimate, vk_mixup, vk_sft_bwd, vk_elbo_denoise, vk_elbo_sisr.
Prints one JSON line per kernel; CUDA events, 20 iterations after 3 warm-ups, working sets larger than L2 where the
training shapes allow it (b = 64 patches of 128^2)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from virnet_b200 import ops  # noqa: E402
from virnet_b200.datasets.DenoisingDatasets import SimulateTrainGPU  # noqa: E402
from virnet_b200.loss.resize_right import downsample_matrix  # noqa: E402
from virnet_b200.utils.util_denoising import gaussian_window  # noqa: E402

dev = "cuda"
PEAK = 6546.0
try:
    PEAK = float(json.load(open(Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"))["hbm_gbps_sustained"])
except Exception:  # noqa: BLE001
    pass


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def report(name, secs, nbytes, note):
    gbs = nbytes / secs / 1e9
    print(json.dumps(dict(kernel=name, us=round(secs * 1e6, 1), algorithmic_MB=round(nbytes / 1e6, 1),
                          achieved_GBps=round(gbs, 1), peak_GBps=PEAK, frac=round(gbs / PEAK, 3), note=note)), flush=True)


n, p = 64, 128
px = n * p * p
patches = torch.randint(0, 256, (n, p, p, 3), dtype=torch.uint8, device=dev)
noise = torch.randn(n, p, p, 3, device=dev)
ds = SimulateTrainGPU(pch_size=p)
params = torch.tensor([ds.draw_sigma_params() for _ in range(n)], dtype=torch.float64, device=dev)
aug = torch.arange(n, dtype=torch.int32, device=dev) % 8
t = timeit(lambda: ops.synth_denoise(patches, params, aug, noise))
report("vk_synth_denoise", t, px * (3 + 12 + 12 + 12 + 4), "b=64 128^2: u8 patch + fp32 noise in, noisy + gt + sigma map out")

im_noisy, im_gt, _ = ops.synth_denoise(patches, params, aug, noise)
win = gaussian_window(7, dev)
out = torch.empty_like(im_noisy)
t = timeit(lambda: ops.noise_estimate(im_noisy, im_gt, win, out))
report("vk_noise_estimate", t, px * 3 * 4 * 3, "b=64 128^2 3ch, 7x7 window: 2 reads + 1 write")

perm = torch.randperm(n, device=dev)
lam = torch.rand(n, device=dev)
t = timeit(lambda: ops.mixup(im_gt, im_noisy, perm, lam))
report("vk_mixup", t, px * 3 * 4 * 6, "b=64 128^2 3ch: 2 tensors x (2 reads + 1 write)")

c = 96
g = torch.randn(16, 256, 256, c, device=dev).to(torch.bfloat16)
x = torch.randn_like(g)
r = torch.randn_like(g)
gx = torch.empty_like(g)
mul = torch.rand(16, c, device=dev)
dm, dd = torch.zeros(16, c, device=dev), torch.zeros(16, c, device=dev)
t = timeit(lambda: ops.sft_bwd(g, x, mul, gx, dm, dd, dtype=ops.VK_BF16, c=c, resid=r))
report("vk_sft_bwd", t, g.numel() * 2 * 4, "b=16 256^2 C=96 bf16 with residual: 3 reads + 1 write")

mu = torch.rand(n, 3, p, p, device=dev)
sg = torch.rand(n, 1, p, p, device=dev) * 0.01 + 1e-4
d_mu, d_sg = torch.empty_like(mu), torch.empty_like(sg)
t = timeit(lambda: ops.elbo_denoise(mu, sg, im_noisy, im_gt, sg, beta0_scale=24.5, eps2=1e-6, alpha0=24.5,
                                    digamma_am1=3.1355727, d_mu=d_mu, d_sigma=d_sg))
report("vk_elbo_denoise", t, px * (44 + 16), "b=64 128^2: 44 B read + 16 B written per pixel")

N, sf, h = 16, 4, 64
H = h * sf
mu = torch.rand(N, 3, H, H, device=dev)
hr = torch.rand_like(mu)
lr = torch.rand(N, 3, h, h, device=dev)
sig = torch.rand(N, device=dev) * 0.01 + 1e-4
kin = torch.stack([1 + torch.rand(N, device=dev), 1 + torch.rand(N, device=dev), torch.rand(N, device=dev) - 0.5], 1).contiguous()
gam = torch._standard_gamma(torch.full((N, 2), 49.0, device=dev))
rho = torch.randn(N, device=dev)
z = torch.randn_like(mu)
rh = downsample_matrix(H, sf, "bicubic", dev)
t = timeit(lambda: ops.elbo_sisr(mu, hr, lr, sig, kin, kin, sig, sig.log(), gam, rho, z, rh, rh, k_size=21, center=10.0,
                                 alpha0=40.5, digamma_am1=3.6788, kappa0=50.0, r2=1e-4, eps2=1e-5, pk0=0.02, pk1=2.0))
fma = N * 3 * (H * H * 441 * 2 + (H + 20) * (H + 20) * 441)
print(json.dumps(dict(kernel="vk_elbo_sisr (13 launches)", us=round(t * 1e6, 1), blur_GFMA=round(fma / 1e9, 2),
                      achieved_TFMAps=round(fma / t / 1e12, 2),
                      note="b=16 x4 64->256: CUDA-core bound (three 21x21 blur passes); HBM traffic ~0.3 GB")), flush=True)
