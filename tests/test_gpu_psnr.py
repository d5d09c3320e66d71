"""PSNR parity with the reference on its own test images (north_star: "PSNR within 0.01 dB"; SURVEY.md §8d
protocol): fixtures hold crops of test_data/CBSD68 and test_data/Set5, the noisy / low-resolution inputs, the
UNMODIFIED reference's outputs with the seed-1234 weights and its PSNR (tools/gen_golden_psnr.py).
PSNR = utils/util_image.py:68-89 on uint8 images.  tf32 mode (the 1e-3 parity mode)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def psnr_uint8(a, b):
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return float("inf") if mse == 0 else 20 * math.log10(255.0 / math.sqrt(mse))


def to_uint8_hwc(mu):
    return (mu[0].clamp(0, 1) * 255.0).round().to(torch.uint8).permute(1, 2, 0).cpu()


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_denoise_psnr_parity_cbsd68(golden_dir):
    import virnet_b200
    fx = torch.load(golden_dir / "psnr_parity.pt")
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=[96, 192, 288], dep_S=5, n_resblocks=3,
                                    noise_cond=True, extra_mode="Input", noise_avg=False, precision="tf32").cuda().eval()
    for case in fx["denoise"]:
        with torch.no_grad():
            mu, _ = net(case["noisy"].cuda())
        assert rel(mu.cpu(), case["mu"]) < 1e-3, case["name"]
        p = psnr_uint8(to_uint8_hwc(mu), case["gt8"])
        assert abs(p - case["psnr"]) <= 0.01, (case["name"], p, case["psnr"])


def test_sisr_psnr_parity_set5(golden_dir):
    import virnet_b200
    fx = torch.load(golden_dir / "psnr_parity.pt")
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNetSR(im_chn=3, sigma_chn=1, dep_S=5, dep_K=8, n_feat=[96, 160, 224], n_resblocks=2,
                                      extra_mode="Both", noise_avg=True, noise_cond=True, kernel_cond=True,
                                      precision="tf32").cuda().eval()
    for case in fx["sisr"]:
        with torch.no_grad():
            mu, kinfo, sigma = net(case["lr"].cuda(), 4)
        assert rel(mu.cpu(), case["mu"]) < 1e-3, case["name"]
        assert rel(kinfo.cpu(), case["kinfo"]) < 1e-3 and rel(sigma.cpu(), case["sigma"]) < 1e-3
        p = psnr_uint8(to_uint8_hwc(mu), case["gt8"])
        assert abs(p - case["psnr"]) <= 0.01, (case["name"], p, case["psnr"])
