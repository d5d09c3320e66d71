"""Host-side random draws of the data-side drop-ins consume Python's `random` / torch's CPU generator in the reference's
order (CPU only: no kernels are launched)."""
import random
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
import gen_golden_synth as GS  # noqa: E402


def test_simulate_train_gpu_draws_follow_reference_order(golden_dir):
    """SimulateTrainGPU.draw_sigma_params + the augmentation flag reproduce what replay_draws (the reference's
    __getitem__ order, datasets/DenoisingDatasets.py:189-209,218-238) yields after the crop draws."""
    from virnet_b200.datasets.DenoisingDatasets import SimulateTrainGPU
    images = [im.numpy() for im in torch.load(golden_dir / "synth_denoise.pt")["images"]]
    for mode in ("niid", "iid"):
        for seed in (3, 21, 40):
            _, params, aug, _ = GS.replay_draws(seed, images, mode)
            random.seed(seed)
            ind = random.randint(0, len(images) - 1)                      # the loader's own draws: image, crop offsets
            random.randint(0, images[ind].shape[0] - GS.PCH)
            random.randint(0, images[ind].shape[1] - GS.PCH)
            ds = SimulateTrainGPU(pch_size=GS.PCH, mode=mode)
            mine = ds.draw_sigma_params()
            # the reference draws the noise tensor from torch's generator here (no `random` consumption)
            assert mine == params and random.randint(0, 7) == aug


def test_mixup_draw_order_matches_reference():
    """datasets/data_tools.py:21-27: randperm first, then the Beta(0.6, 0.6) rsample, both on the CPU generator."""
    from virnet_b200.datasets.data_tools import MixUp_AUG
    torch.manual_seed(17)
    idx, lam = MixUp_AUG().draw(6, "cpu")
    torch.manual_seed(17)
    dist = torch.distributions.beta.Beta(torch.tensor([0.6]), torch.tensor([0.6]))
    exp_idx = torch.randperm(6)
    exp_lam = dist.rsample((6, 1)).view(-1)
    assert torch.equal(idx, exp_idx) and torch.equal(lam, exp_lam)


def test_data_side_dropins_refuse_cpu_tensors():
    """No CPU fallback: the drop-ins raise on CPU tensors instead of computing something else."""
    import pytest
    from virnet_b200.datasets.data_tools import MixUp_AUG
    from virnet_b200.datasets.DenoisingDatasets import SimulateTrainGPU
    from virnet_b200.datasets.SISRDatasets import GeneralTrainGPU
    from virnet_b200.utils.util_denoising import noise_estimate_fun
    x = torch.rand(2, 3, 32, 32)
    with pytest.raises(RuntimeError):
        noise_estimate_fun(x, x, 7)
    with pytest.raises(RuntimeError):
        MixUp_AUG().aug(x, x)
    with pytest.raises(RuntimeError):
        SimulateTrainGPU(32).synthesize(torch.zeros(2, 32, 32, 3, dtype=torch.uint8))
    with pytest.raises(RuntimeError):
        GeneralTrainGPU(4).degrade(torch.rand(1, 3, 48, 48))


def test_adam_bias_corrections_match_the_kernel_formula():
    """trainer._bias_corrections feeds the CUDA-graph path (vk_adam_clip_step_dev) with exactly what vk_adam_clip_step
    computes for the eager path: double arithmetic on the betas rounded to fp32, results rounded to fp32."""
    import numpy as np
    from virnet_b200.trainer import _bias_corrections
    for betas in ((0.9, 0.999), (0.5, 0.9), (0.95, 0.98)):
        for step in (1, 2, 10, 1000, 123456):
            b1, b2 = (float(np.float32(b)) for b in betas)
            want = (1.0 - b1 ** step, (1.0 - b2 ** step) ** 0.5)
            got = _bias_corrections(betas, step)
            assert np.float32(got[0]) == np.float32(want[0]) and np.float32(got[1]) == np.float32(want[1])
            assert 0.0 < got[0] <= 1.0 and 0.0 < got[1] <= 1.0
