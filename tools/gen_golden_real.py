"""Golden vectors for the real-noise trainer's data-side operators: runs the UNMODIFIED
reference utils/util_denoising.py:noise_estimate_fun (OpenCV Gaussian window) on CPU in the build container.
   python tools/gen_golden_real.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import ref_import  # noqa: E402

OUT = ROOT / "tests" / "golden"
CASES = {"k7_3ch": (2, 3, 40, 52, 7), "k9_1ch_ragged": (1, 1, 33, 37, 9), "k7_tiny": (1, 3, 5, 6, 7)}


def real_inputs(n, c, h, w, seed=31):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(n, c, h, w, generator=g)
    noisy = gt + 0.1 * torch.rand(n, 1, h, w, generator=g) * torch.randn(n, c, h, w, generator=g)
    return noisy, gt


def main():
    ref_import.import_reference()
    import importlib
    ud = importlib.import_module("utils.util_denoising")
    out = {}
    for name, (n, c, h, w, k) in CASES.items():
        noisy, gt = real_inputs(n, c, h, w)
        out[name] = ud.noise_estimate_fun(noisy, gt, k).clone()
        print(name, out[name].mean().item())
    torch.save(out, OUT / "noise_estimate.pt")


if __name__ == "__main__":
    main()
