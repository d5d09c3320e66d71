"""Timed comparators for bench.py — NOT product code, never imported by virnet_b200/.

Two things are timed beside the CUDA path, on the same box:

* the reference's own CPU PyTorch path on the host cores (`bench.py --impl reference`, and the bounded
  `cpu_baseline` leg of the default run);
* the "kernel to beat": the same reference module in eager PyTorch / cuDNN on the same B200, fp32 NCHW with
  PyTorch's default TF32 convolutions and bf16 autocast + channels_last (`gpu_comparator` in the bench line).

`kind` says what ran: "reference" = the UNMODIFIED reference modules (`/root/reference`, or its staged copy
`baseline/_ref/` made by baseline/install_ref.py — git-ignored, it travels to the GPU box with the snapshot);
"port" = oracle/virnet_oracle.py, the functional restatement that tests pin bit-exactly on the reference (the
same ATen / cuDNN / oneDNN calls, so it times the same kernels).  The training step is the loop body of
train_denoising_syn.py:175-184: zero_grad, forward, elbo_denoising_simple, backward, clip_grad_norm_ on RNet / SNet
parameters, Adam.step (the reference's four `.item()` host syncs per step are left out: they would only slow it).
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_FEAT, N_RES, DEP_S = [96, 192, 288], 3, 5
ALPHA0, EPS2 = 24.5, 1e-6
SR_KW = dict(im_chn=3, sigma_chn=1, dep_S=5, dep_K=8, n_feat=[96, 160, 224], n_resblocks=2, extra_mode="Both",
             noise_avg=True, noise_cond=True, kernel_cond=True)


def reference_modules():
    """(networks.VIRNet, loss.ELBO_simple) of the unmodified reference, or None."""
    try:
        import ref_import
        if not ref_import.available():
            return None
        return ref_import.import_reference()
    except Exception as e:  # noqa: BLE001
        print(f"comparator: reference import failed ({e!r}); using the oracle port", file=sys.stderr)
        return None


class DenoiseStep:
    """One training step of the denoising-syn network with stock PyTorch ops on `device`."""

    def __init__(self, device, b, batch, mode="fp32", lr=1e-4, clip_R=1e3, clip_S=1e2):
        import torch
        from torch import nn
        self.torch, self.nn = torch, nn
        self.device, self.mode = torch.device(device), mode
        self.autocast = mode == "bf16_autocast_channels_last"
        self.clip_R, self.clip_S = clip_R, clip_S
        mods = reference_modules()
        torch.manual_seed(1234)                                     # train_denoising_syn.py:52-53
        if mods is not None:
            vir, elbo = mods
            self.kind = "reference"
            net = vir.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=N_FEAT, dep_S=DEP_S, n_resblocks=N_RES,
                                    noise_cond=True, extra_mode="Input", noise_avg=False).to(self.device)
            if self.autocast:
                net = net.to(memory_format=torch.channels_last)
            net.train()
            self.net, self.loss_fn = net, elbo.elbo_denoising_simple
            named = list(net.named_parameters())
            self.fwd = net
        else:
            from oracle import virnet_oracle as O
            self.kind = "port"
            cfg = O.NetCfg(n_feat=tuple(N_FEAT), n_resblocks=N_RES, dep_S=DEP_S)
            sd = O.build_state_dict(cfg)
            params = {}
            for k, v in sd.items():
                v = v.to(self.device)
                if self.autocast and v.dim() == 4:
                    v = v.contiguous(memory_format=torch.channels_last)
                params[k] = v.clone().requires_grad_(True)
            self.loss_fn = O.elbo_denoising_simple
            named = list(params.items())
            self.fwd = lambda x: O.vir_denoise_forward(params, x, cfg)
        self.params = [p for _, p in named]
        self.pR = [p for n, p in named if "rnet" in n.lower()]      # train_denoising_syn.py:153-154
        self.pS = [p for n, p in named if "snet" in n.lower()]
        self.opt = torch.optim.Adam(self.params, lr=lr)
        fmt = torch.channels_last if self.autocast else torch.contiguous_format
        self.im_noisy, self.im_gt, sigma_gt = [t.to(self.device).contiguous(memory_format=fmt) for t in batch]
        # train_denoising_syn.py:157,172: alpha0 is a 1-element tensor on the device
        self.alpha0 = torch.tensor([ALPHA0], dtype=torch.float32, device=self.device)
        self.beta0 = self.alpha0 * sigma_gt
        self.b = b

    def __call__(self):
        torch, nn = self.torch, self.nn
        self.opt.zero_grad()
        if self.autocast:
            with torch.autocast(self.device.type, dtype=torch.bfloat16):
                mu, sigma = self.fwd(self.im_noisy)
            mu, sigma = mu.float(), sigma.float()
        else:
            mu, sigma = self.fwd(self.im_noisy)
        loss, *_ = self.loss_fn(mu, sigma, self.im_noisy, self.im_gt, EPS2, self.alpha0, self.beta0)
        loss.backward()
        nn.utils.clip_grad_norm_(self.pR, self.clip_R)
        nn.utils.clip_grad_norm_(self.pS, self.clip_S)
        self.opt.step()
        return loss


def cpu_train(batch, b, steps, warmup, budget_s=None, min_steps=5):
    """patches/s of the reference's CPU path on all host cores:
    (value, sec_per_step, cores, kind, timed steps, (min, median, max) s/step).  With `budget_s`, timing stops early
    once the wall clock since the start exceeds it (but not before `min_steps` timed steps)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = DenoiseStep("cpu", b, batch, mode="fp32")
    times = []
    t_start = time.perf_counter()
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        t1 = time.perf_counter()
        if it >= warmup:
            times.append(t1 - t0)
            if budget_s is not None and len(times) >= min(min_steps, steps) and t1 - t_start > budget_s:
                break
    n = len(times)
    sec = sum(times) / n
    times.sort()
    return b / sec, sec, cores, step.kind, n, (times[0], times[n // 2], times[-1])


def gpu_train(device, batch, b, steps, warmup, mode):
    """patches/s of the stock-PyTorch step on the GPU (CUDA events, after warm-up)."""
    import torch
    torch.backends.cudnn.benchmark = True          # give cuDNN its autotuned best
    step = DenoiseStep(device, b, batch, mode=mode)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    return {"value": b / (ms * 1e-3), "unit": "patches/s", "ms_per_step": ms, "batch": b, "steps": steps,
            "warmup": warmup, "kind": step.kind, "mode": mode}


def gpu_infer(device, x, steps, warmup, mode, sr_sf=0):
    """img/s of the stock-PyTorch forward (eval, no_grad): denoising net, or the SISR net when sr_sf > 0."""
    import torch
    torch.backends.cudnn.benchmark = True
    dev = torch.device(device)
    autocast = mode == "bf16_autocast_channels_last"
    mods = reference_modules()
    torch.manual_seed(1234)
    if mods is not None:
        vir, _ = mods
        kind = "reference"
        if sr_sf:
            net = vir.VIRAttResUNetSR(**SR_KW)
            fwd = lambda t: net(t, sr_sf)
        else:
            net = vir.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=N_FEAT, dep_S=DEP_S, n_resblocks=N_RES,
                                    noise_cond=True, extra_mode="Input", noise_avg=False)
            fwd = net
        net = net.to(dev).eval()
        if autocast:
            net = net.to(memory_format=torch.channels_last)
    else:
        from oracle import virnet_oracle as O
        kind = "port"
        if sr_sf:
            cfg = O.NetCfg(im_chn=3, sigma_chn=1, n_feat=(96, 160, 224), n_resblocks=2, dep_S=5, dep_K=8,
                           extra_mode="Both", noise_avg=True, sisr=True)
        else:
            cfg = O.NetCfg(n_feat=tuple(N_FEAT), n_resblocks=N_RES, dep_S=DEP_S)
        sd = {k: v.to(dev) for k, v in O.build_state_dict(cfg).items()}
        fwd = (lambda t: O.vir_sisr_forward(sd, t, sr_sf, cfg)) if sr_sf else (lambda t: O.vir_denoise_forward(sd, t, cfg))
    x = x.to(dev)
    if autocast:
        x = x.contiguous(memory_format=torch.channels_last)

    def run():
        with torch.no_grad():
            if autocast:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return fwd(x)
            return fwd(x)

    for _ in range(warmup):
        run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        run()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    return {"value": x.shape[0] / (ms * 1e-3), "unit": "img/s", "ms_per_step": ms, "batch": x.shape[0],
            "steps": steps, "warmup": warmup, "kind": kind, "mode": mode}
