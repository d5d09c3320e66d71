"""PSNR-parity fixtures (north_star: "PSNR within 0.01 dB"): run the UNMODIFIED reference on crops of its own
test images (test_data/CBSD68, test_data/Set5) with the seed-1234 initial weights and store inputs, outputs
and the reference's PSNR.  model_zoo/ is empty and there is no network, so the weights are the initial ones
(outputs are not good restorations; the test checks that both implementations produce the SAME image).

    python tools/gen_golden_psnr.py      # needs /root/reference and cv2; writes tests/golden/psnr_parity.pt
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT))

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ref_import  # noqa: E402

REF = Path("/root/reference")
OUT = ROOT / "tests" / "golden" / "psnr_parity.pt"


def psnr_uint8(a, b):
    """utils/util_image.py:68-89 calculate_psnr (border 0): uint8 images, float64 mse."""
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else float(20 * np.log10(255.0 / np.sqrt(mse)))


def to_uint8(t):
    return (t.clamp(0, 1) * 255.0).round().to(torch.uint8)


def main():
    vir, _ = ref_import.import_reference()
    torch.set_num_threads(8)
    rng = np.random.default_rng(1000)                       # scripts/denoising_virnet_syn.py:96
    out = {"denoise": [], "sisr": []}

    torch.manual_seed(1234)
    net = vir.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=[96, 192, 288], dep_S=5, n_resblocks=3, noise_cond=True,
                            extra_mode="Input", noise_avg=False).eval()
    for name, (y0, x0, hh, ww) in (("101085.png", (40, 60, 81, 98)), ("102061.png", (100, 30, 64, 64))):
        im = cv2.imread(str(REF / "test_data" / "CBSD68" / name), cv2.IMREAD_COLOR)[:, :, ::-1]
        gt = np.ascontiguousarray(im[y0:y0 + hh, x0:x0 + ww]).astype(np.float32) / 255.0
        noisy = gt + rng.standard_normal(gt.shape).astype(np.float32) * (25.0 / 255.0)
        x = torch.from_numpy(noisy.transpose(2, 0, 1)).unsqueeze(0).contiguous()
        with torch.no_grad():
            mu, _ = net(x)
        mu8 = to_uint8(mu[0]).permute(1, 2, 0).numpy()
        gt8 = (gt * 255.0).round().astype(np.uint8)
        out["denoise"].append(dict(name=name, noisy=x.clone(), gt8=torch.from_numpy(gt8), mu=mu.clone(),
                                   psnr=psnr_uint8(mu8, gt8)))
        print(name, mu.shape, out["denoise"][-1]["psnr"])

    torch.manual_seed(1234)
    netsr = vir.VIRAttResUNetSR(im_chn=3, sigma_chn=1, dep_S=5, dep_K=8, n_feat=[96, 160, 224], n_resblocks=2,
                                extra_mode="Both", noise_avg=True, noise_cond=True, kernel_cond=True).eval()
    for name, (y0, x0, hh, ww) in (("butterfly_GT.bmp", (64, 64, 128, 96)),):
        im = cv2.imread(str(REF / "test_data" / "Set5" / name), cv2.IMREAD_COLOR)[:, :, ::-1]
        gt = np.ascontiguousarray(im[y0:y0 + hh, x0:x0 + ww]).astype(np.float32) / 255.0
        lr = cv2.resize(gt, (ww // 4, hh // 4), interpolation=cv2.INTER_CUBIC)
        lr = lr + rng.standard_normal(lr.shape).astype(np.float32) * (2.55 / 255.0)
        x = torch.from_numpy(np.ascontiguousarray(lr.transpose(2, 0, 1))).unsqueeze(0).contiguous()
        with torch.no_grad():
            mu, kinfo, sigma = netsr(x, 4)
        mu8 = to_uint8(mu[0]).permute(1, 2, 0).numpy()
        gt8 = (gt * 255.0).round().astype(np.uint8)
        out["sisr"].append(dict(name=name, lr=x.clone(), gt8=torch.from_numpy(gt8), mu=mu.clone(), kinfo=kinfo.clone(),
                                sigma=sigma.clone(), psnr=psnr_uint8(mu8, gt8)))
        print(name, mu.shape, out["sisr"][-1]["psnr"])
    torch.save(out, OUT)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
