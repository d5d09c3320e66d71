#!/usr/bin/env python
"""Bitwise repeatability of the super-resolution training step (SISRTrainer.step with fixed draws and lr = 0): which
engine buffers differ between repeated runs.  Forward buffers must never differ; gradient buffers downstream of the
small per-sample kernels that accumulate with atomics (KNet / SFT parameter gradients) may."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import virnet_b200  # noqa: E402
from virnet_b200.trainer import SISRTrainer  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    dev = torch.device("cuda")
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNetSR(im_chn=3, sigma_chn=1, kernel_chn=3, n_feat=[96, 160, 224], dep_S=5, dep_K=8,
                                      noise_cond=True, kernel_cond=True, n_resblocks=2, extra_mode="Both",
                                      noise_avg=True, precision=prec).to(dev)
    net.engine().deterministic = True
    lr_sz, sf = 64, 4
    g = torch.Generator(device=dev).manual_seed(0)
    im_hr = torch.rand(B, 3, lr_sz * sf, lr_sz * sf, device=dev, generator=g)
    im_lr = torch.nn.functional.avg_pool2d(im_hr, sf) + 0.01 * torch.randn(B, 3, lr_sz, lr_sz, device=dev, generator=g)
    kinfo_gt = torch.stack([0.5 + 3 * torch.rand(B, device=dev, generator=g), 0.5 + 3 * torch.rand(B, device=dev, generator=g),
                            torch.rand(B, device=dev, generator=g) - 0.5], dim=1)
    nlevel = torch.full((B, 1, 1, 1), (2.55 / 255) ** 2, device=dev)
    tr = SISRTrainer(net, sf)
    eng = tr.engine
    draws = tr._draw(im_hr)
    ref = None
    for it in range(iters):
        terms = tr.step(im_hr, im_lr, kinfo_gt, nlevel, lr=0.0, draws=draws)
        torch.cuda.synchronize()
        snap = {"terms": terms.clone(), "mu": tr.last_mu.clone(), "kinfo": tr.last_kinfo.clone(), "sigma": tr.last_sigma.clone(),
                "flat_grads": eng.flat_grads.clone()}
        for key, st in eng._sets.items():
            for (name, shape, dt), t in st["bufs"].items():
                snap[f"buf:{name}"] = t.clone()
        names = [n for n, _ in net.named_parameters()]
        if ref is None:
            ref = snap
            print(f"run 0: {len(snap)} tensors recorded, loss {terms[0].item():.6f}", flush=True)
            continue
        bad = []
        for k, v in snap.items():
            r = ref[k]
            if not torch.equal(v.view(torch.uint8), r.view(torch.uint8)):
                idx = (v.view(torch.uint8) != r.view(torch.uint8)).flatten().nonzero().flatten()
                bad.append((k, tuple(v.shape), int(idx.numel())))
        print(f"run {it}: {len(bad)} tensors differ", flush=True)
        for rec in bad[:60]:
            print("   ", rec, flush=True)
        if any(k == "flat_grads" for k, _, _ in bad):
            d = (snap["flat_grads"] != ref["flat_grads"])
            offs = eng.flat_offsets + [eng.flat_total]
            hit = [names[i] for i in range(len(names)) if d[offs[i]:offs[i + 1]].any()]
            print(f"    parameters whose gradient differs: {len(hit)} of {len(names)}: {hit[:12]} ...", flush=True)


if __name__ == "__main__":
    main()
