"""Pins the oracle's synthetic-batch restatement (oracle/virnet_oracle.py:synth_denoise_sample, data_aug_np) against the
unmodified reference SimulateTrain.__getitem__ (tests/golden/synth_denoise.pt, tools/gen_golden_synth.py): the draws are
replayed from Python's `random` and torch's generator in the reference's order."""
import sys
from pathlib import Path

import pytest
import torch

from oracle import virnet_oracle as O

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
import gen_golden_synth as G  # noqa: E402

CASES = [(m, s) for m, seeds in G.SEEDS.items() for s in seeds]


@pytest.mark.parametrize("mode,seed", CASES)
def test_oracle_synth_matches_reference(mode, seed, golden_dir):
    gold = torch.load(golden_dir / "synth_denoise.pt")
    images = [im.numpy() for im in gold["images"]]
    patch, params, aug, noise = G.replay_draws(seed, images, mode)
    im_noisy, im_gt, sigma_gt = O.synth_denoise_sample(patch, params, aug, noise, clip=mode.endswith("clip"))
    r_noisy, r_gt, r_sigma = gold["samples"][(mode, seed)]
    assert torch.equal(im_gt, r_gt)
    torch.testing.assert_close(sigma_gt, r_sigma, rtol=1e-6, atol=1e-12)
    torch.testing.assert_close(im_noisy, r_noisy, rtol=1e-6, atol=1e-7)


def test_golden_covers_all_augmentations(golden_dir):
    gold = torch.load(golden_dir / "synth_denoise.pt")
    images = [im.numpy() for im in gold["images"]]
    flags = {G.replay_draws(s, images, m)[2] for m, s in CASES}
    assert len(flags) >= 6, flags
