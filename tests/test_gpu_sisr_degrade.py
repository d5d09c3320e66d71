"""SISR training-pair synthesis on the B200 (SISR half): vk_sisr_degrade through
virnet_b200.datasets.SISRDatasets.GeneralTrainGPU against the reference's outputs (tests/golden/sisr_degrade.pt) and the
oracle at training size.  fp32 on the device vs float64 intermediate in scipy / numpy: 1e-5."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
import gen_golden_sisr_degrade as G  # noqa: E402


@pytest.mark.parametrize("name", list(G.CASES))
def test_degrade_vs_reference_golden(name, golden_dir):
    from virnet_b200.datasets.SISRDatasets import GeneralTrainGPU
    ref = torch.load(golden_dir / "sisr_degrade.pt")[name]
    sf, h, w, ds, shift, seed = G.CASES[name]
    hr = torch.from_numpy(G.hr_patch(h, w, seed)).permute(2, 0, 1)[None].contiguous().cuda()
    gen = GeneralTrainGPU(sf, kernel_shift=shift, downsampler=ds)
    _, im_lr, im_blur, _, nlevel = gen.degrade(hr, kernels=ref["kernel"][None].float(), std=torch.tensor([ref["std"]]),
                                               noise=ref["noise"].permute(2, 0, 1)[None].contiguous())
    torch.testing.assert_close(im_blur[0].permute(1, 2, 0).cpu(), ref["im_blur"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(im_lr[0].permute(1, 2, 0).cpu(), ref["im_lr"], rtol=1e-5, atol=1e-5)
    assert nlevel.shape == (1, 1, 1, 1)


def test_degrade_training_size_batch_vs_oracle():
    import random
    from oracle import virnet_oracle as O
    from virnet_b200.datasets.SISRDatasets import GeneralTrainGPU
    sf, n, H = 4, 3, 192
    rng = np.random.default_rng(3)
    hr = rng.random((n, H, H, 3), dtype=np.float32)
    gen = GeneralTrainGPU(sf)
    random.seed(12)
    ks, _, sd = zip(*[gen.draw() for _ in range(n)])
    noise = torch.randn(n, 3, H // sf, H // sf, generator=torch.Generator().manual_seed(1))
    _, im_lr, im_blur, _, _ = gen.degrade(torch.from_numpy(hr).permute(0, 3, 1, 2).contiguous().cuda(),
                                          kernels=torch.from_numpy(np.stack(ks)).float(), std=torch.tensor(sd),
                                          noise=noise)
    for k in range(n):
        o_blur, o_lr = O.sisr_degrade_sample(hr[k], ks[k], sf, "Bicubic", noise[k].permute(1, 2, 0).numpy(), sd[k])
        torch.testing.assert_close(im_blur[k].permute(1, 2, 0).cpu(), torch.from_numpy(o_blur), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(im_lr[k].permute(1, 2, 0).cpu(), torch.from_numpy(o_lr), rtol=1e-5, atol=1e-5)
    assert im_lr.min().item() >= 0.0 and im_lr.max().item() <= 1.0
