"""oracle/eval_protocol.py (metrics, noise / degradation recipes of the reference's evaluation scripts) pinned on
values the UNMODIFIED reference produced (tests/golden/eval_kat.json, tools/gen_golden_eval.py), and the oracle
network at FULL image size against the reference's outputs.  CPU only."""
import numpy as np
import pytest
import torch

from eval_common import cbsd68_images, checksum_close, kat, set5_images
from oracle import eval_protocol as E
from oracle import virnet_oracle as O


@pytest.fixture(scope="module")
def K():
    return kat()


@pytest.fixture(scope="module")
def noisy(K):
    return E.niid_noisy_images(cbsd68_images(K))


def test_niid_noise_recipe_reproduces_the_reference_inputs(K, noisy):
    """scripts/denoising_virnet_syn.py:96-131: same rng stream, same nearest-exact resize of the `peaks` map."""
    for x, e in zip(noisy, K["denoise"]):
        assert list(x.shape[:2]) == e["shape"] and x.dtype == np.float32
        checksum_close(x, e["input"], K["stride"], 1e-9)


def test_metrics_match_the_reference_known_answers(K, noisy):
    """calculate_psnr (bit-exact) and calculate_ssim (1e-12) incl. the Y channel and border cropping."""
    for x, gt, e in zip(noisy, cbsd68_images(K), K["denoise"]):
        n8 = E.img_as_ubyte(np.clip(x, 0.0, 1.0))
        m = e["metric_kat"]
        assert E.calculate_psnr(n8, gt, 0, False) == m["psnr_rgb"]
        assert E.calculate_psnr(n8, gt, 4, True) == m["psnr_y_b4"]
        assert abs(E.calculate_ssim(n8, gt, 0, False) - m["ssim_rgb"]) < 1e-12
        assert abs(E.calculate_ssim(n8, gt, 16, True) - m["ssim_y_b16"]) < 1e-12


def test_sisr_degradation_recipe_reproduces_the_reference_inputs(K):
    """scripts/sisr_virnet_syn.py:105-141 / util_sisr.degrade_virnet(nlevel=2.55, seed=1234, Bicubic)."""
    kernel, _ = E.shifted_anisotropic_gaussian(21, 4, (0.6 * 4) ** 2, (0.6 * 4) ** 2, 0, False)
    checksum_close(kernel, K["sisr_kernel"], K["stride"], 1e-12)
    for im, e in zip(set5_images(K), K["sisr"]):
        gt = E.modcrop(im, 4)
        assert list(gt.shape[:2]) == e["shape"]
        lr = E.degrade_virnet(gt.astype(np.float32) / 255.0, kernel, 4)
        checksum_close(lr, e["input"], K["stride"], 2e-6)


def test_oracle_network_at_full_image_size_matches_the_reference(K, noisy):
    """One full CBSD68 image (481x321 -> padded 484x324) and one Set5 image x4: mu sample, PSNR, SSIM."""
    torch.set_num_threads(min(8, torch.get_num_threads()))
    cfg = O.NetCfg(n_feat=(96, 192, 288), n_resblocks=3, dep_S=5)
    torch.manual_seed(1234)
    sd = O.build_state_dict(cfg)
    e, x, gt = K["denoise"][0], noisy[0], cbsd68_images(K)[0]
    with torch.no_grad():
        mu, _ = O.vir_denoise_forward(sd, torch.from_numpy(x.transpose(2, 0, 1)[None]), cfg)
    checksum_close(mu.numpy(), e["mu"], K["stride"], 1e-5)
    den8 = E.img_as_ubyte(mu.clamp(0, 1)[0].numpy().transpose(1, 2, 0))
    assert abs(E.calculate_psnr(den8, gt) - e["psnr"]) < 1e-3 and abs(E.calculate_ssim(den8, gt) - e["ssim"]) < 1e-5

    cfg = O.NetCfg(n_feat=(96, 160, 224), n_resblocks=2, dep_S=5, dep_K=8, extra_mode="Both", noise_avg=True, sisr=True)
    torch.manual_seed(1234)
    sd = O.build_state_dict(cfg)
    i = [s["name"] for s in K["sisr"]].index("butterfly_GT")
    e, gt = K["sisr"][i], E.modcrop(set5_images(K)[i], 4)
    kernel, _ = E.shifted_anisotropic_gaussian(21, 4, (0.6 * 4) ** 2, (0.6 * 4) ** 2, 0, False)
    lr = E.degrade_virnet(gt.astype(np.float32) / 255.0, kernel, 4)
    with torch.no_grad():
        mu, kinfo, sigma = O.vir_sisr_forward(sd, torch.from_numpy(lr.transpose(2, 0, 1)[None]), 4, cfg)
    checksum_close(mu.numpy(), e["mu"], K["stride"], 1e-4)
    sr8 = E.img_as_ubyte(mu.clamp(0, 1)[0].numpy().transpose(1, 2, 0))
    assert abs(E.calculate_psnr(sr8, gt, 16, True) - e["psnr_y"]) < 1e-3
    assert abs(E.calculate_ssim(sr8, gt, 16, True) - e["ssim_y"]) < 1e-5
