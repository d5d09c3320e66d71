"""Super-resolution forward path on the B200: the drop-in VIRAttResUNetSR
(KNet + SFT-modulated RNet through the C ABI) against the reference-generated fixture (kat.json /
sisr_x4_48_slices.pt, tools/gen_golden.py) and against the CPU oracle on the same weights.

Tolerances: 1e-3 relative in "tf32" mode (north_star's forward bar), 1e-2 in "bf16"."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(n_feat=(96, 160, 224), n_resblocks=2)


def make_sr(precision, n_feat=CFG["n_feat"], n_res=CFG["n_resblocks"], dep_K=8):
    import virnet_b200
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNetSR(im_chn=3, sigma_chn=1, dep_S=5, dep_K=dep_K, n_feat=list(n_feat),
                                      n_resblocks=n_res, extra_mode="Both", noise_avg=True, noise_cond=True,
                                      kernel_cond=True, precision=precision)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return net.cuda().eval(), sd


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("precision,tol", [("tf32", 1e-3), ("bf16", 1e-2)])
def test_sisr_x4_forward_vs_reference_fixture(precision, tol, kat, golden_dir):
    net, _ = make_sr(precision)
    x = torch.rand(2, 3, 48, 48, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        mu, kinfo, sigma = net(x.cuda(), 4)
    k = kat["sisr_x4_48"]
    assert mu.shape == (2, 3, 192, 192) and kinfo.shape == (2, 3) and sigma.shape == (2, 1, 1, 1)
    torch.testing.assert_close(kinfo.cpu(), torch.tensor(k["kinfo"]), rtol=tol * 5, atol=tol)
    torch.testing.assert_close(sigma.flatten().cpu(), torch.tensor(k["sigma"]), rtol=tol * 5, atol=tol)
    assert abs(mu.double().mean().item() - k["mu"]["mean"]) < tol * abs(k["mu"]["mean"]) * 5
    sl = torch.load(golden_dir / "sisr_x4_48_slices.pt")
    assert rel(mu[0, :, :8, :8].cpu(), sl["mu_slice"]) < tol * 5


@pytest.mark.parametrize("shape,sf", [((1, 3, 21, 30), 4), ((2, 3, 24, 24), 2)])
def test_sisr_forward_vs_oracle_ragged(shape, sf):
    """Odd LR sizes (reflect pad of the upsampled grid, crop) and another scale factor, small net."""
    from oracle import virnet_oracle as O
    net, sd = make_sr("tf32", n_feat=(32, 64, 96), n_res=2, dep_K=3)
    cfg = O.NetCfg(n_feat=(32, 64, 96), n_resblocks=2, extra_mode="Both", noise_avg=True, sisr=True, dep_K=3)
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        mu, kinfo, sigma = net(x.cuda(), sf)
        mu_o, kinfo_o, sigma_o = O.vir_sisr_forward(sd, x, sf, cfg)
    assert mu.shape == mu_o.shape
    assert rel(kinfo.cpu(), kinfo_o) < 1e-3 and rel(sigma.cpu(), sigma_o) < 1e-3
    assert rel(mu.cpu(), mu_o) < 1e-3


def test_sisr_training_mode_raises():
    net, _ = make_sr("tf32", n_feat=(32, 64, 96), n_res=1, dep_K=2)
    net.train()
    with pytest.raises(NotImplementedError):
        net(torch.rand(1, 3, 16, 16, device="cuda"), 4)
