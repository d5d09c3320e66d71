"""SISR negative ELBO on the B200: virnet_b200.loss.ELBO_simple.elbo_sisr through the
C ABI (vk_elbo_sisr) against (1) the reference's own outputs (tests/golden/sisr_loss.pt, produced by the
unmodified reference with the same seeded draws) and (2) the CPU oracle with autograd on larger / ragged cases.

Tolerances: everything is fp32 on both sides; loss terms 1e-4 relative, gradients 1e-3 relative L2 (the
reference reduces in a different order and inverts the 2x2 covariance with LU)."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
import gen_golden_sisr_loss as G  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def run_cuda(inputs, sf, ds, shift, draws, hyper=G.HYPER):
    from virnet_b200.loss.ELBO_simple import elbo_sisr
    mu, sigma_est, kinfo_est, im_hr, im_lr, sigma_prior, kinfo_gt = [t.detach().cuda() for t in inputs]
    for t in (mu, sigma_est, kinfo_est):
        t.requires_grad_(True)
    loss, detail = elbo_sisr(mu=mu, sigma_est=sigma_est, kinfo_est=kinfo_est, im_hr=im_hr, im_lr=im_lr,
                             sigma_prior=sigma_prior, alpha0=torch.tensor([hyper["alpha0"]]).cuda(), kinfo_gt=kinfo_gt,
                             kappa0=torch.tensor([hyper["kappa0"]]).cuda(), r2=hyper["r2"], eps2=hyper["eps2"], sf=sf,
                             k_size=hyper["k_size"], penalty_K=hyper["penalty_K"], shift=shift, downsampler=ds,
                             draws=tuple(d.cuda() for d in draws))
    loss.backward()
    return loss, detail, mu.grad, kinfo_est.grad, sigma_est.grad


@pytest.mark.parametrize("name", list(G.CASES))
def test_elbo_sisr_vs_reference_golden(name, golden_dir):
    from oracle import virnet_oracle as O
    ref = torch.load(golden_dir / "sisr_loss.pt")[name]
    n, h, w, sf, ds, shift = G.CASES[name]
    inputs = G.sisr_loss_inputs(n, h, w, sf)
    torch.manual_seed(G.DRAW_SEED)
    draws = O.reference_draws(n, inputs[0].shape, G.HYPER["kappa0"])        # CPU generator: what the fixture consumed
    loss, detail, d_mu, d_kinfo, d_sigma = run_cuda(inputs, sf, ds, shift, draws)
    torch.testing.assert_close(loss.detach().cpu().reshape(()), ref["loss"].reshape(()), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(torch.stack([d.detach().cpu() for d in detail[:7]]), ref["terms"].reshape(7), rtol=1e-4,
                               atol=1e-4)
    assert detail[7].shape == ref["kernel"].shape
    assert rel(detail[7].cpu(), ref["kernel"]) < 1e-4
    assert rel(d_mu.cpu(), ref["d_mu"]) < 1e-3
    assert rel(d_kinfo.cpu(), ref["d_kinfo"]) < 1e-3
    assert rel(d_sigma.cpu(), ref["d_sigma"]) < 1e-3


@pytest.mark.parametrize("n,h,w,sf,ds,shift", [(3, 48, 48, 4, "Bicubic", False), (2, 37, 50, 2, "Bicubic", True),
                                               (2, 33, 20, 3, "Direct", False), (1, 6, 7, 4, "Bicubic", True)])
def test_elbo_sisr_vs_oracle(n, h, w, sf, ds, shift):
    """Training-size patch (48 -> 192), ragged sizes spanning several 32x32 tiles, and an image barely larger
    than the blur kernel's reflect padding."""
    from oracle import virnet_oracle as O
    inputs = G.sisr_loss_inputs(n, h, w, sf, seed=23)
    torch.manual_seed(99)
    draws = O.reference_draws(n, inputs[0].shape, G.HYPER["kappa0"])
    mu, sigma_est, kinfo_est, im_hr, im_lr, sigma_prior, kinfo_gt = [t.clone() for t in inputs]
    for t in (mu, sigma_est, kinfo_est):
        t.requires_grad_(True)
    lo, do = O.elbo_sisr(mu, sigma_est, kinfo_est, im_hr, im_lr, sigma_prior, G.HYPER["alpha0"], kinfo_gt,
                         G.HYPER["kappa0"], G.HYPER["r2"], G.HYPER["eps2"], sf, G.HYPER["k_size"], G.HYPER["penalty_K"],
                         shift, ds, gamma_draw=draws[0], rho_draw=draws[1], z_draw=draws[2])
    lo.backward()
    loss, detail, d_mu, d_kinfo, d_sigma = run_cuda(inputs, sf, ds, shift, draws)
    torch.testing.assert_close(loss.detach().cpu().reshape(()), lo.detach().reshape(()), rtol=1e-4, atol=1e-4)
    for a, b in zip(detail[:7], do[:7]):
        torch.testing.assert_close(a.detach().cpu().reshape(()), b.detach().reshape(()), rtol=1e-4, atol=1e-4)
    assert rel(detail[7].cpu(), do[7].detach()) < 1e-4
    assert rel(d_mu.cpu(), mu.grad) < 1e-3
    assert rel(d_kinfo.cpu(), kinfo_est.grad) < 1e-3
    assert rel(d_sigma.cpu(), sigma_est.grad) < 1e-3


def test_elbo_sisr_rho_clamp_and_prior_map():
    """rho pushed outside [-1, 1] (clamped: no gradient through the correlation draw) and a spatial prior map."""
    from oracle import virnet_oracle as O
    n, h, w, sf = 2, 12, 12, 4
    inputs = list(G.sisr_loss_inputs(n, h, w, sf, seed=5))
    inputs[2] = inputs[2].clone()
    inputs[2][0, 2] = 0.9999                       # + sqrt(r2) * draw crosses 1 for a positive draw
    inputs[5] = (1e-4 + 1e-2 * torch.rand(n, 1, h, w, generator=torch.Generator().manual_seed(1)))
    torch.manual_seed(3)
    draws = list(O.reference_draws(n, inputs[0].shape, G.HYPER["kappa0"]))
    draws[1] = draws[1].abs() + 0.5
    mu, sigma_est, kinfo_est, im_hr, im_lr, sigma_prior, kinfo_gt = [t.clone() for t in inputs]
    for t in (mu, sigma_est, kinfo_est):
        t.requires_grad_(True)
    lo, do = O.elbo_sisr(mu, sigma_est, kinfo_est, im_hr, im_lr, sigma_prior, G.HYPER["alpha0"], kinfo_gt,
                         G.HYPER["kappa0"], G.HYPER["r2"], G.HYPER["eps2"], sf, 21, G.HYPER["penalty_K"], False,
                         "Bicubic", gamma_draw=draws[0], rho_draw=draws[1], z_draw=draws[2])
    lo.backward()
    loss, detail, d_mu, d_kinfo, d_sigma = run_cuda(inputs, sf, "Bicubic", False, draws)
    torch.testing.assert_close(loss.detach().cpu().reshape(()), lo.detach().reshape(()), rtol=1e-4, atol=1e-4)
    assert rel(d_kinfo.cpu(), kinfo_est.grad) < 1e-3 and rel(d_sigma.cpu(), sigma_est.grad) < 1e-3


def test_elbo_sisr_internal_draws_are_seeded():
    """Without explicit draws the loss consumes torch's CUDA generator: same seed -> same value."""
    from virnet_b200.loss.ELBO_simple import elbo_sisr
    inputs = [t.cuda() for t in G.sisr_loss_inputs(2, 12, 12, 4)]
    mu, sigma_est, kinfo_est, im_hr, im_lr, sigma_prior, kinfo_gt = inputs
    vals = []
    for seed in (7, 7, 8):
        torch.manual_seed(seed)
        loss, _ = elbo_sisr(mu, sigma_est, kinfo_est, im_hr, im_lr, sigma_prior, 40.5, kinfo_gt, 50.0, 1e-4, 1e-5, 4, 21,
                            [0.02, 2], False, "Bicubic")
        vals.append(loss.item())
    assert vals[0] == vals[1] and vals[0] != vals[2]
