"""Golden vectors for the SISR training-pair synthesis (SISR half): composes the UNMODIFIED
reference functions exactly as datasets/SISRDatasets.py:78-104 does (util_sisr.shifted_anisotropic_Gaussian,
util_sisr.imconv_np, ResizeRight.resize, Gaussian noise + clips) on seeded HR patches.
   python tools/gen_golden_sisr_degrade.py"""
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

OUT = ROOT / "tests" / "golden"
CASES = {"x4_bicubic": (4, 48, 56, "Bicubic", False, 5), "x2_bicubic_shift": (2, 40, 34, "Bicubic", True, 6),
         "x3_direct": (3, 45, 39, "Direct", False, 7)}


def hr_patch(h, w, seed):
    return np.random.default_rng(seed).random((h, w, 3), dtype=np.float32)


def draws(sf, seed):
    random.seed(seed)
    torch.manual_seed(seed)
    lam1 = random.uniform(0.2, sf)
    lam2 = random.uniform(lam1, sf) if random.random() < 0.7 else lam1
    theta = random.uniform(0, np.pi)
    return lam1, lam2, theta


def main():
    import ref_import
    ref_import.import_reference()
    import importlib
    us = importlib.import_module("utils.util_sisr")
    from ResizeRight.resize_right import resize
    out = {}
    for name, (sf, h, w, ds, shift, seed) in CASES.items():
        im_hr = hr_patch(h, w, seed)
        lam1, lam2, theta = draws(sf, seed)
        kernel, infos = us.shifted_anisotropic_Gaussian(k_size=21, sf=sf, lambda_1=lam1 ** 2, lambda_2=lam2 ** 2,
                                                        theta=theta, shift=shift)
        im_blur = us.imconv_np(im_hr, kernel, padding_mode="reflect", correlate=False)
        im_blur = np.clip(im_blur, a_min=0.0, a_max=1.0)
        if ds.lower() == "direct":
            im_blur = im_blur[::sf, ::sf, ]
        else:
            im_blur = resize(im_blur, scale_factors=1 / sf).astype(np.float32)
        std = random.uniform(0.1, 15) / 255.0
        noise = torch.randn(im_blur.shape, dtype=torch.float32).numpy()
        im_lr = np.clip(im_blur + noise * std, a_min=0, a_max=1.0)
        out[name] = dict(kernel=torch.from_numpy(np.ascontiguousarray(kernel)), infos=torch.from_numpy(infos),
                         im_blur=torch.from_numpy(np.ascontiguousarray(im_blur)).float(),
                         im_lr=torch.from_numpy(np.ascontiguousarray(im_lr)).float(), noise=torch.from_numpy(noise),
                         std=float(std))
        print(name, im_blur.shape, float(im_lr.mean()))
    torch.save(out, OUT / "sisr_degrade.pt")


if __name__ == "__main__":
    main()
