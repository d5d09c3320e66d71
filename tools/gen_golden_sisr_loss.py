"""Golden vectors for the SISR loss: runs the UNMODIFIED reference
loss/ELBO_simple.py:elbo_sisr (with utils/util_sisr.py and ResizeRight) on CPU in the build container and
stores loss terms, the blur kernel and the gradients w.r.t. (mu, kinfo_est, sigma_est).

Inputs and the loss's internal random draws are reproducible from the seeds below: the reference consumes
torch's global generator in the order Gamma.rsample -> randn_like(rho) -> randn_like(mu), which
oracle.virnet_oracle.reference_draws replays.   python tools/gen_golden_sisr_loss.py
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import ref_import  # noqa: E402

OUT = ROOT / "tests" / "golden"

CASES = {
    # name: (n, lr_h, lr_w, sf, downsampler, shift)
    "x4_bicubic": (2, 12, 12, 4, "Bicubic", False),
    "x2_bicubic_ragged": (2, 15, 13, 2, "Bicubic", True),
    "x3_bicubic": (1, 11, 16, 3, "Bicubic", False),
    "x4_direct": (2, 12, 9, 4, "Direct", True),
}
HYPER = dict(alpha0=0.5 * 9 ** 2, kappa0=50.0, r2=1e-4, eps2=1e-5, k_size=21, penalty_K=[0.02, 2])
DRAW_SEED = 4321


def sisr_loss_inputs(n, h, w, sf, seed=11):
    g = torch.Generator().manual_seed(seed)
    im_hr = torch.rand(n, 3, h * sf, w * sf, generator=g)
    mu = (im_hr + 0.05 * torch.randn(n, 3, h * sf, w * sf, generator=g)).clone()
    im_lr = torch.rand(n, 3, h, w, generator=g)
    kinfo_est = torch.stack([0.5 + 4 * torch.rand(n, generator=g), 0.5 + 4 * torch.rand(n, generator=g),
                             torch.rand(n, generator=g) * 1.6 - 0.8], dim=1)
    kinfo_gt = torch.stack([0.5 + 4 * torch.rand(n, generator=g), 0.5 + 4 * torch.rand(n, generator=g),
                            torch.rand(n, generator=g) * 1.6 - 0.8], dim=1)
    sigma_est = (1e-4 + 1e-2 * torch.rand(n, 1, 1, 1, generator=g))
    sigma_prior = (1e-4 + 1e-2 * torch.rand(n, 1, 1, 1, generator=g))
    return mu, sigma_est, kinfo_est, im_hr, im_lr, sigma_prior, kinfo_gt


def main():
    _, elbo = ref_import.import_reference()
    out = {}
    for name, (n, h, w, sf, ds, shift) in CASES.items():
        mu, sigma_est, kinfo_est, im_hr, im_lr, sigma_prior, kinfo_gt = sisr_loss_inputs(n, h, w, sf)
        for t in (mu, sigma_est, kinfo_est):
            t.requires_grad_(True)
        torch.manual_seed(DRAW_SEED)
        loss, detail = elbo.elbo_sisr(mu=mu, sigma_est=sigma_est, kinfo_est=kinfo_est, im_hr=im_hr, im_lr=im_lr,
                                      sigma_prior=sigma_prior, alpha0=torch.tensor([HYPER["alpha0"]]), kinfo_gt=kinfo_gt,
                                      kappa0=torch.tensor([HYPER["kappa0"]]), r2=HYPER["r2"], eps2=HYPER["eps2"], sf=sf,
                                      k_size=HYPER["k_size"], penalty_K=HYPER["penalty_K"], shift=shift,
                                      downsampler=ds)
        loss.backward()
        out[name] = dict(loss=loss.detach().clone(), terms=torch.stack([d.detach() for d in detail[:7]]),
                         kernel=detail[7].detach().clone(), d_mu=mu.grad.clone(), d_kinfo=kinfo_est.grad.clone(),
                         d_sigma=sigma_est.grad.clone())
        print(name, float(loss), [round(float(d), 5) for d in detail[:7]])
    torch.save(out, OUT / "sisr_loss.pt")


if __name__ == "__main__":
    main()
