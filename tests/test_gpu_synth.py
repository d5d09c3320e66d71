"""Device-side synthesis of denoising training batches: vk_synth_denoise through
virnet_b200.datasets.DenoisingDatasets.SimulateTrainGPU against the reference's own samples
(tests/golden/synth_denoise.pt) and the oracle.  fp32 both sides: clean image bit-exact, maps / noisy 1e-6."""
import random
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
import gen_golden_synth as G  # noqa: E402

CASES = [(m, s) for m, seeds in G.SEEDS.items() for s in seeds]


def test_synth_batch_vs_reference_golden(golden_dir):
    """All golden samples of one mode as ONE batch (mixed augmentation flags and centres in a launch)."""
    from virnet_b200.datasets.DenoisingDatasets import SimulateTrainGPU
    gold = torch.load(golden_dir / "synth_denoise.pt")
    images = [im.numpy() for im in gold["images"]]
    for mode, seeds in G.SEEDS.items():
        draws = [G.replay_draws(s, images, mode) for s in seeds]
        patches = torch.stack([torch.from_numpy(d[0].copy()) for d in draws]).cuda()
        params = torch.tensor([d[1] for d in draws], dtype=torch.float64)
        aug = torch.tensor([d[2] for d in draws], dtype=torch.int32)
        noise = torch.stack([torch.from_numpy(d[3]) for d in draws]).cuda()
        ds = SimulateTrainGPU(pch_size=G.PCH, chn=3, mode=mode.split("_")[0], clip=mode.endswith("clip"))
        im_noisy, im_gt, sigma_gt = ds.synthesize(patches, params=params, aug=aug, noise=noise)
        for k, s in enumerate(seeds):
            r_noisy, r_gt, r_sigma = gold["samples"][(mode, s)]
            assert torch.equal(im_gt[k].cpu(), r_gt), (mode, s)
            torch.testing.assert_close(sigma_gt[k].cpu(), r_sigma, rtol=1e-6, atol=1e-12)
            torch.testing.assert_close(im_noisy[k].cpu(), r_noisy, rtol=1e-6, atol=1e-7)


def test_synth_full_size_vs_oracle_and_properties():
    from oracle import virnet_oracle as O
    from virnet_b200.datasets.DenoisingDatasets import SimulateTrainGPU
    n, p = 8, 128
    g = torch.Generator().manual_seed(1)
    patches = torch.randint(0, 256, (n, p, p, 3), generator=g, dtype=torch.uint8)
    noise = torch.randn(n, p, p, 3, generator=g)
    ds = SimulateTrainGPU(pch_size=p, mode="niid")
    random.seed(5)
    rows, flags = [], []
    for k in range(n):
        rows.append(ds.draw_sigma_params())
        flags.append(k % 8)                      # every augmentation once
    # centres outside / on the border of the patch exercise the analytic min / max of the bump
    rows[0][0], rows[0][1] = 0.0, 127.9
    rows[1][0], rows[1][1] = 63.5, 0.2
    im_noisy, im_gt, sigma_gt = ds.synthesize(patches.cuda(), params=torch.tensor(rows, dtype=torch.float64),
                                              aug=torch.tensor(flags, dtype=torch.int32), noise=noise.cuda())
    for k in range(n):
        o_noisy, o_gt, o_sigma = O.synth_denoise_sample(patches[k].numpy(), rows[k], flags[k], noise[k].numpy())
        assert torch.equal(im_gt[k].cpu(), o_gt)
        torch.testing.assert_close(sigma_gt[k].cpu(), o_sigma, rtol=1e-6, atol=1e-12)
        torch.testing.assert_close(im_noisy[k].cpu(), o_noisy, rtol=1e-6, atol=1e-7)
        # the map spans exactly [down^2, up^2]
        assert abs(sigma_gt[k].max().item() - rows[k][3] ** 2) < 1e-6 and abs(sigma_gt[k].min().item() - max(rows[k][4] ** 2, 1e-10)) < 1e-6


def test_synth_internal_draws_follow_python_random():
    from virnet_b200.datasets.DenoisingDatasets import SimulateTrainGPU
    ds = SimulateTrainGPU(pch_size=32, mode="niid")
    patches = torch.randint(0, 256, (4, 32, 32, 3), dtype=torch.uint8).cuda()
    outs = []
    for seed in (9, 9, 10):
        random.seed(seed)
        torch.manual_seed(seed)
        outs.append(ds.synthesize(patches)[2].clone())
    assert torch.equal(outs[0], outs[1]) and not torch.equal(outs[0], outs[2])
