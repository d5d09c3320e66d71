"""Pins the oracle's SISR training-pair synthesis (oracle/virnet_oracle.py:sisr_degrade_sample) and the product's host-side
kernel generator against the unmodified reference functions composed as datasets/SISRDatasets.py:78-104
(tests/golden/sisr_degrade.pt, tools/gen_golden_sisr_degrade.py)."""
import random
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import virnet_oracle as O

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
import gen_golden_sisr_degrade as G  # noqa: E402


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_degrade_matches_reference(name, golden_dir):
    ref = torch.load(golden_dir / "sisr_degrade.pt")[name]
    sf, h, w, ds, shift, seed = G.CASES[name]
    im_blur, im_lr = O.sisr_degrade_sample(G.hr_patch(h, w, seed), ref["kernel"].numpy(), sf, ds, ref["noise"].numpy(),
                                           ref["std"])
    torch.testing.assert_close(torch.from_numpy(im_blur), ref["im_blur"], rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(torch.from_numpy(im_lr), ref["im_lr"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", list(G.CASES))
def test_product_kernel_generator_matches_reference(name, golden_dir):
    """virnet_b200.datasets.SISRDatasets draws (lam1, lam2, theta, std) in the reference's order and builds its kernel."""
    from virnet_b200.datasets.SISRDatasets import GeneralTrainGPU
    ref = torch.load(golden_dir / "sisr_degrade.pt")[name]
    sf, h, w, ds, shift, seed = G.CASES[name]
    random.seed(seed)
    kernel, infos, std = GeneralTrainGPU(sf, kernel_shift=shift, downsampler=ds).draw()
    np.testing.assert_allclose(kernel, ref["kernel"].numpy(), rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(infos, ref["infos"].numpy(), rtol=1e-10)
    assert abs(std - ref["std"]) < 1e-15
