"""Pins the oracle's noise_estimate_fun / Gaussian window against the unmodified reference (OpenCV window;
tests/golden/noise_estimate.pt from tools/gen_golden_real.py) and checks the product's host-side window."""
import sys
from pathlib import Path

import pytest
import torch

from oracle import virnet_oracle as O

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
import gen_golden_real as G  # noqa: E402


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_noise_estimate_matches_reference(name, golden_dir):
    ref = torch.load(golden_dir / "noise_estimate.pt")[name]
    n, c, h, w, k = G.CASES[name]
    noisy, gt = G.real_inputs(n, c, h, w)
    torch.testing.assert_close(O.noise_estimate_fun(noisy, gt, k), ref, rtol=1e-5, atol=1e-9)


def test_product_window_equals_oracle_window():
    from virnet_b200.utils.util_denoising import gaussian_window
    for k in (5, 7, 9):
        assert torch.equal(gaussian_window(k, "cpu"), O.inverse_gamma_window(k))
        assert abs(gaussian_window(k, "cpu").sum().item() - 1.0) < 1e-6
