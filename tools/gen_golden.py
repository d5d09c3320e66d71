"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through tools/ref_import.py) on CPU in the build container.

The reference ships no tests or golden vectors; these fixtures are what pins
the oracle (oracle/virnet_oracle.py) and, through it, the CUDA path.  Re-run with
    python tools/gen_golden.py
whenever the set of cases changes.  Inputs are seeded, so only outputs are stored.
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import ref_import  # noqa: E402

OUT = ROOT / "tests" / "golden"


def den_inputs(n, h, w, seed=0):
    g = torch.Generator().manual_seed(seed)
    im_gt = torch.rand(n, 3, h, w, generator=g)
    sig = 5 / 255 + torch.rand(n, 1, h, w, generator=g) * 70 / 255
    im_noisy = im_gt + torch.randn(n, 3, h, w, generator=g) * sig
    sigma_gt = (sig ** 2).clamp_min(1e-10)
    return im_noisy, im_gt, sigma_gt


def stats(t):
    return dict(mean=t.double().mean().item(), std=t.double().std().item(), absmax=t.abs().max().item())


def run_denoise(vir, elbo, n_feat, n_res, n, h, w, backward=True):
    torch.manual_seed(1234)
    net = vir.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=n_feat, dep_S=5, n_resblocks=n_res, noise_cond=True,
                            extra_mode="Input", noise_avg=False)
    im_noisy, im_gt, sigma_gt = den_inputs(n, h, w)
    alpha0 = torch.tensor([24.5])
    out = {}
    out["param_sum"] = sum(p.double().sum().item() for p in net.parameters())
    out["param_abs_sum"] = sum(p.double().abs().sum().item() for p in net.parameters())
    out["param_count"] = sum(p.numel() for p in net.parameters())
    mu, sigma = net(im_noisy)
    out["mu"], out["sigma"] = stats(mu), stats(sigma)
    tensors = {"mu": mu.detach(), "sigma": sigma.detach()}
    if backward:
        loss, lh, kg, ig = elbo.elbo_denoising_simple(mu, sigma, im_noisy, im_gt, 1e-6, alpha0, alpha0 * sigma_gt)
        loss.backward()
        out["loss"] = dict(loss=loss.item(), lh=lh.item(), kl_gauss=kg.item(), kl_igamma=ig.item())
        gr = {k: p.grad for k, p in net.named_parameters()}
        out["grad_norm_RNet"] = sum(v.double().pow(2).sum().item() for k, v in gr.items() if "rnet" in k.lower()) ** 0.5
        out["grad_norm_SNet"] = sum(v.double().pow(2).sum().item() for k, v in gr.items() if "snet" in k.lower()) ** 0.5
        out["grad_norms"] = {k: v.double().norm().item() for k, v in gr.items()}
        tensors["grads"] = gr
    return out, tensors


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    vir, elbo = ref_import.import_reference()
    torch.set_num_threads(8)
    kat = {"torch": torch.__version__, "reference": "zsyOAOA/VIRNet @ e3d1934 (imported from /root/reference)"}

    # (a) the BASELINE denoising config, 2 x 128 x 128 (statistics + small slices)
    o, t = run_denoise(vir, elbo, [96, 192, 288], 3, 2, 128, 128)
    kat["den_syn_128"] = o
    torch.save({"mu_slice": t["mu"][0, :, :8, :8].clone(), "sigma_slice": t["sigma"][0, :, :8, :8].clone(),
                "grad_tail_w": t["grads"]["RNet.tail.weight"].clone(),
                "grad_snet_last_w": t["grads"]["SNet.conv_last.weight"].clone()}, OUT / "den_syn_128_slices.pt")
    # (b) odd size -> exercises reflect padding + crop
    o, t = run_denoise(vir, elbo, [96, 192, 288], 3, 1, 37, 50, backward=False)
    kat["den_syn_37x50"] = o
    torch.save({"mu": t["mu"].clone(), "sigma": t["sigma"].clone()}, OUT / "den_syn_37x50.pt")
    # (c) a small net whose complete outputs and gradients fit in a fixture
    o, t = run_denoise(vir, elbo, [32, 64, 96], 2, 2, 32, 32)
    kat["den_small_32"] = o
    torch.save({"mu": t["mu"], "sigma": t["sigma"],
                "grads": {k: v.clone() for k, v in t["grads"].items() if v.numel() <= 4096}}, OUT / "den_small_32.pt")
    # (d) small net, odd size, with gradients (reflect-pad backward)
    o, t = run_denoise(vir, elbo, [32, 64, 96], 2, 1, 21, 27)
    kat["den_small_21x27"] = o
    torch.save({"mu": t["mu"], "sigma": t["sigma"]}, OUT / "den_small_21x27.pt")

    # (e) SISR forward
    torch.manual_seed(1234)
    netsr = vir.VIRAttResUNetSR(im_chn=3, sigma_chn=1, dep_S=5, dep_K=8, n_feat=[96, 160, 224], n_resblocks=2,
                                extra_mode="Both", noise_avg=True, noise_cond=True, kernel_cond=True)
    x = torch.rand(2, 3, 48, 48, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        mu, kinfo, sigma = netsr(x, 4)
    kat["sisr_x4_48"] = dict(mu=stats(mu), kinfo=kinfo.tolist(), sigma=sigma.flatten().tolist(),
                             param_sum=sum(p.double().sum().item() for p in netsr.parameters()),
                             param_count=sum(p.numel() for p in netsr.parameters()))
    torch.save({"mu_slice": mu[0, :, :8, :8].clone()}, OUT / "sisr_x4_48_slices.pt")

    # (f) the loss alone on random tensors
    g = torch.Generator().manual_seed(7)
    mu = torch.rand(2, 3, 16, 16, generator=g).requires_grad_(True)
    sg = (torch.rand(2, 1, 16, 16, generator=g) * 0.05 + 1e-3).requires_grad_(True)
    y = torch.rand(2, 3, 16, 16, generator=g)
    gt = torch.rand(2, 3, 16, 16, generator=g)
    b0 = 24.5 * (torch.rand(2, 1, 16, 16, generator=g) * 0.05 + 1e-3)
    loss, lh, kg, ig = elbo.elbo_denoising_simple(mu, sg, y, gt, 1e-6, torch.tensor([24.5]), b0)
    loss.backward()
    kat["elbo_16"] = dict(loss=loss.item(), lh=lh.item(), kl_gauss=kg.item(), kl_igamma=ig.item())
    torch.save({"d_mu": mu.grad.clone(), "d_sigma": sg.grad.clone()}, OUT / "elbo_16.pt")

    (OUT / "kat.json").write_text(json.dumps(kat, indent=1, sort_keys=True))
    print(json.dumps({k: v for k, v in kat.items() if k != "den_syn_128"}, indent=1)[:1500])
    print("den_syn_128:", {k: v for k, v in kat["den_syn_128"].items() if k != "grad_norms"})


if __name__ == "__main__":
    main()
