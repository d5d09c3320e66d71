"""Real-noise trainer's data-side operators on the B200: vk_noise_estimate and vk_mixup
through the drop-ins virnet_b200.utils.util_denoising.noise_estimate_fun / datasets.data_tools.MixUp_AUG against the
reference's outputs (tests/golden/noise_estimate.pt) and the oracle; DenoiseTrainer.step_real against the manual
composition.  Tolerance: fp32 both sides, 1e-5 relative."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
import gen_golden_real as G  # noqa: E402


@pytest.mark.parametrize("name", list(G.CASES))
def test_noise_estimate_vs_reference_golden(name, golden_dir):
    from virnet_b200.utils.util_denoising import noise_estimate_fun
    ref = torch.load(golden_dir / "noise_estimate.pt")[name]
    n, c, h, w, k = G.CASES[name]
    noisy, gt = G.real_inputs(n, c, h, w)
    got = noise_estimate_fun(noisy.cuda(), gt.cuda(), k).cpu()
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-9)


def test_noise_estimate_full_size_vs_oracle():
    from oracle import virnet_oracle as O
    from virnet_b200.utils.util_denoising import noise_estimate_fun
    noisy, gt = G.real_inputs(4, 3, 128, 128, seed=2)
    got = noise_estimate_fun(noisy.cuda(), gt.cuda(), 7).cpu()
    torch.testing.assert_close(got, O.noise_estimate_fun(noisy, gt, 7), rtol=1e-5, atol=1e-9)
    assert got.min().item() >= 1e-10


def test_mixup_matches_reference_draws_and_blend():
    from oracle import virnet_oracle as O
    from virnet_b200.datasets.data_tools import MixUp_AUG
    noisy, gt = G.real_inputs(6, 3, 32, 36, seed=4)
    torch.manual_seed(21)
    mine = MixUp_AUG()
    got_gt, got_noisy = mine.aug(gt.cuda(), noisy.cuda())
    # the reference's draw order on the CPU generator (datasets/data_tools.py:22-27)
    torch.manual_seed(21)
    dist = torch.distributions.beta.Beta(torch.tensor([0.6]), torch.tensor([0.6]))
    indices = torch.randperm(6)
    lam = dist.rsample((6, 1)).view(-1)
    exp_gt, exp_noisy = O.mixup(gt, noisy, indices, lam)
    torch.testing.assert_close(got_gt.cpu(), exp_gt, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(got_noisy.cpu(), exp_noisy, rtol=1e-6, atol=1e-7)


def test_trainer_step_real_equals_manual_composition():
    import virnet_b200
    from virnet_b200.datasets.data_tools import MixUp_AUG
    from virnet_b200.trainer import DenoiseTrainer
    from virnet_b200.utils.util_denoising import noise_estimate_fun

    def build():
        torch.manual_seed(1234)
        net = virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=3, n_feat=[32, 64, 96, 128], dep_S=4, n_resblocks=1,
                                        noise_cond=True, extra_mode="Input", noise_avg=False, precision="tf32").cuda()
        return net, DenoiseTrainer(net, lr=1e-4, clip_grad_R=5e2, clip_grad_S=1e2, deterministic=True)

    noisy, gt = [t.cuda() for t in G.real_inputs(4, 3, 32, 32, seed=8)]
    net_a, tr_a = build()
    torch.manual_seed(3)
    la = tr_a.step_real(noisy, gt, var_window=7, mixup=MixUp_AUG()).clone()
    net_b, tr_b = build()
    torch.manual_seed(3)
    g2, n2 = MixUp_AUG().aug(gt, noisy)
    lb = tr_b.step(n2, g2, noise_estimate_fun(n2, g2, 7)).clone()
    assert torch.equal(la, lb)
    # deterministic trainers (ordered split-K reduction, no atomics anywhere on the step): the two compositions must
    # agree bit for bit, not just up to summation order
    assert torch.equal(tr_a.grad_norms, tr_b.grad_norms)
    for pa, pb in zip(net_a.parameters(), net_b.parameters()):
        assert torch.equal(pa, pb)
