"""Fixtures of the FULL PSNR-parity protocol, produced by the UNMODIFIED reference:

* tests/golden/images/: the first 8 CBSD68 images (sorted, 321x481 / 481x321) and the 5 Set5 images, re-encoded
  losslessly as PNG (test data of the reference repository, not source);
* tests/golden/eval_kat.json: for every image
    - checksums of the network INPUT the reference's script builds (scripts/denoising_virnet_syn.py:96-131 with
      rng(1000) and the `peaks` niid map; scripts/sisr_virnet_syn.py:105-141 with degrade_virnet(nlevel=2.55, seed=1234),
      kernel shifted_anisotropic_Gaussian(21, 4, (0.6*4)^2, (0.6*4)^2, 0, False), Bicubic),
    - the reference network's output with the seed-1234 initial weights: mean / std / a strided sample of mu, and the
      PSNR / SSIM the reference's util_image functions report for it,
    - PSNR / SSIM known answers of util_image on (noisy, clean) uint8 pairs (RGB and Y channel, with border).

    python tools/gen_golden_eval.py     # needs /root/reference and cv2
"""
import json
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
IMG_DIR = ROOT / "tests" / "golden" / "images"
OUT = ROOT / "tests" / "golden" / "eval_kat.json"
N_CBSD = 8
STRIDE = 997


def checksum(a):
    a = np.asarray(a, dtype=np.float64)
    return {"sum": float(a.sum()), "sumsq": float((a * a).sum()), "sample": [float(v) for v in a.flatten()[::STRIDE][:64]]}


def main():
    vir, _ = ref_import.import_reference()
    from utils import util_denoising, util_image, util_sisr
    torch.set_num_threads(8)
    IMG_DIR.mkdir(parents=True, exist_ok=True)
    kat = {"stride": STRIDE, "denoise": [], "sisr": []}

    # ---------------- denoising: scripts/denoising_virnet_syn.py ----------------
    rng = util_denoising.noise_generator()
    var_maps = [util_denoising.peaks(256), util_denoising.sincos_kernel(),
                util_denoising.generate_gauss_kernel_mix(256, 256, rng)]
    sigma_max, sigma_min = 75 / 255.0, 10 / 255.0
    sigma_base = var_maps[0]
    sigma_base = sigma_min + (sigma_base - sigma_base.min()) / (sigma_base.max() - sigma_base.min()) * (sigma_max - sigma_min)
    torch.manual_seed(1234)
    net = vir.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=[96, 192, 288], dep_S=5, n_resblocks=3, noise_cond=True,
                            extra_mode="Input", noise_avg=False).eval()
    im_list = sorted(str(x) for x in (REF / "test_data" / "CBSD68").glob("*.png"))[:N_CBSD]
    for im_path in im_list:
        name = Path(im_path).stem
        im_gt = util_image.imread(im_path, chn="rgb", dtype="uint8")
        cv2.imwrite(str(IMG_DIR / f"cbsd68_{name}.png"), im_gt[:, :, ::-1])
        h, w = im_gt.shape[:2]
        sigma = cv2.resize(sigma_base, (w, h), interpolation=cv2.INTER_NEAREST_EXACT).astype(np.float32)
        noise = rng.standard_normal(size=im_gt.shape) * sigma[:, :, np.newaxis]
        im_noisy = im_gt.astype(np.float32) / 255.0 + noise.astype(np.float32)          # img_as_float32(uint8)
        inputs = torch.from_numpy(im_noisy.transpose(2, 0, 1)[np.newaxis,])
        with torch.no_grad():
            mu, sig = net(inputs)
        im_den = util_image.img_as_ubyte(mu.clamp(0.0, 1.0).squeeze(0).numpy().transpose(1, 2, 0)) \
            if hasattr(util_image, "img_as_ubyte") else None
        from skimage import img_as_ubyte
        im_den = img_as_ubyte(mu.clamp(0.0, 1.0).squeeze(0).numpy().transpose(1, 2, 0))
        noisy8 = img_as_ubyte(np.clip(im_noisy, 0.0, 1.0))
        kat["denoise"].append({
            "name": name, "shape": [h, w], "input": checksum(im_noisy), "mu": checksum(mu.numpy()),
            "mu_mean": float(mu.mean()), "mu_std": float(mu.std()), "sigma_mean": float(sig.mean()),
            "psnr": util_image.calculate_psnr(im_den, im_gt, 0, False), "ssim": util_image.calculate_ssim(im_den, im_gt, 0, False),
            "metric_kat": {"psnr_rgb": util_image.calculate_psnr(noisy8, im_gt, 0, False),
                           "psnr_y_b4": util_image.calculate_psnr(noisy8, im_gt, 4, True),
                           "ssim_rgb": util_image.calculate_ssim(noisy8, im_gt, 0, False),
                           "ssim_y_b16": util_image.calculate_ssim(noisy8, im_gt, 16, True)}})
        print(name, (h, w), kat["denoise"][-1]["psnr"], kat["denoise"][-1]["ssim"])

    # ---------------- super-resolution: scripts/sisr_virnet_syn.py ----------------
    sf, p = 4, 21
    kernel = util_sisr.shifted_anisotropic_Gaussian(p, sf, (0.60 * sf) ** 2, (0.60 * sf) ** 2, 0, False)[0]
    kat["sisr_kernel"] = checksum(kernel)
    torch.manual_seed(1234)
    netsr = vir.VIRAttResUNetSR(im_chn=3, sigma_chn=1, dep_S=5, dep_K=8, n_feat=[96, 160, 224], n_resblocks=2,
                                extra_mode="Both", noise_avg=True, noise_cond=True, kernel_cond=True).eval()
    for im_path in sorted((REF / "test_data" / "Set5").glob("*.bmp")):
        name = im_path.stem
        im_gt = util_image.imread(im_path, chn="rgb", dtype="uint8")
        cv2.imwrite(str(IMG_DIR / f"set5_{name}.png"), im_gt[:, :, ::-1])
        im_gt = util_sisr.modcrop(im_gt, sf)
        im_lr = util_sisr.degrade_virnet(im_gt.astype(np.float32) / 255.0, kernel=kernel, sf=sf, nlevel=2.55, qf=None,
                                         downsampler="Bicubic")
        inputs = torch.from_numpy(im_lr.transpose((2, 0, 1))[np.newaxis,]).type(torch.float32)
        with torch.no_grad():
            mu, kinfo, sig = netsr(inputs, sf)
        from skimage import img_as_ubyte
        im_sr = img_as_ubyte(mu.clamp(0.0, 1.0).squeeze(0).numpy().transpose((1, 2, 0)))
        kat["sisr"].append({
            "name": name, "shape": list(im_gt.shape[:2]), "input": checksum(im_lr), "mu": checksum(mu.numpy()),
            "mu_mean": float(mu.mean()), "mu_std": float(mu.std()), "kinfo": kinfo[0].tolist(), "sigma": float(sig.flatten()[0]),
            "psnr_y": util_image.calculate_psnr(im_sr, im_gt, sf ** 2, True),
            "ssim_y": util_image.calculate_ssim(im_sr, im_gt, sf ** 2, True)})
        print(name, im_gt.shape, kat["sisr"][-1]["psnr_y"], kat["sisr"][-1]["ssim_y"])
    OUT.write_text(json.dumps(kat, indent=1))
    print("wrote", OUT, OUT.stat().st_size, "bytes;", sum(f.stat().st_size for f in IMG_DIR.glob("*.png")), "bytes of images")


if __name__ == "__main__":
    main()
