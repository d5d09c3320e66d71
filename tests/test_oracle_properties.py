"""Size-independent properties of the oracle's restatements (CPU): they hold for the reference by construction and are
what the full-size GPU parity tests lean on where no golden vector exists."""
import math

import numpy as np
import torch

from oracle import virnet_oracle as O


def test_augmentations_form_the_dihedral_group():
    """utils/util_image.py:391-436: the 8 modes are distinct, each is a bijection of the pixels, and mode 4 (rot180)
    is an involution — the kernels' source-index tables must satisfy the same."""
    img = np.arange(5 * 5 * 2, dtype=np.float32).reshape(5, 5, 2)
    outs = [O.data_aug_np(img, m) for m in range(8)]
    for a in range(8):
        assert sorted(outs[a].ravel().tolist()) == sorted(img.ravel().tolist())
        for b in range(a + 1, 8):
            assert not np.array_equal(outs[a], outs[b])
    assert np.array_equal(O.data_aug_np(outs[4], 4), img)
    assert np.array_equal(O.data_aug_np(outs[1], 1), img)
    assert np.array_equal(O.data_aug_np(O.data_aug_np(img, 2), 6), img)          # rot90 then rot270


def test_resize_matrix_rows_are_a_partition_of_unity_and_local():
    for in_sz, sf in ((64, 2), (96, 3), (192, 4), (50, 4)):
        m = O.resize_matrix(in_sz, sf, "bicubic")
        torch.testing.assert_close(m.sum(1), torch.ones(m.shape[0]), rtol=0, atol=1e-5)
        for y in range(m.shape[0]):
            nz = torch.nonzero(m[y]).flatten()
            assert nz.max() - nz.min() < 4 * sf + 1                               # antialiased cubic support
        # a constant image stays constant, a linear ramp is reproduced away from the mirrored borders
        ramp = torch.arange(in_sz, dtype=torch.float32)
        out = m @ ramp
        centres = torch.arange(m.shape[0]) * sf + (in_sz - 1) / 2 - (m.shape[0] - 1) * sf / 2   # resize_right.py:251-262
        k = 3
        torch.testing.assert_close(out[k:-k], centres[k:-k].float(), rtol=0, atol=1e-3)


def test_sigma2kernel_is_a_normalised_centred_gaussian():
    cov = torch.tensor([[[[4.0, 1.0], [1.0, 2.0]]], [[[0.5, 0.0], [0.0, 9.0]]]])
    for shift, sf in ((False, 4), (True, 2)):
        k = O.sigma2kernel(cov, 21, sf, shift)
        assert k.shape == (2, 1, 21, 21)
        torch.testing.assert_close(k.sum((1, 2, 3)), torch.ones(2), rtol=0, atol=1e-6)
        centre = 10 + 0.5 * (sf - 1) if shift else 10
        ax = torch.arange(21, dtype=torch.float32)
        mx = (k[:, 0].sum(2) * ax).sum(1)                                        # centre of mass along the row index
        my = (k[:, 0].sum(1) * ax).sum(1)
        torch.testing.assert_close(mx, torch.full((2,), float(centre)), rtol=0, atol=0.05)
        torch.testing.assert_close(my, torch.full((2,), float(centre)), rtol=0, atol=0.05)


def test_blur_downsample_is_linear_and_mass_preserving():
    g = torch.Generator().manual_seed(0)
    a, b = torch.rand(1, 3, 24, 24, generator=g), torch.rand(1, 3, 24, 24, generator=g)
    k = O.sigma2kernel(torch.tensor([[[[2.0, 0.3], [0.3, 1.0]]]]), 21, 2, False)
    f = lambda x: O.blur_downsample(x, k, 2, "bicubic")
    torch.testing.assert_close(f(2 * a - 3 * b), 2 * f(a) - 3 * f(b), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(f(torch.ones(1, 3, 24, 24)), torch.ones(1, 3, 12, 12), rtol=0, atol=1e-5)


def test_elbo_sisr_gradient_of_clamped_rho_is_the_kl_term_only():
    """ELBO_simple.py:76-77: when rho + sqrt(r2) eps leaves [-1, 1] the likelihood stops depending on kinfo[:, 2]; what
    remains is the Gaussian KL gradient (kinfo - gt) / r2 * penalty_K[0] * penalty_K[1] / 3 / N."""
    n, sf = 2, 2
    g = torch.Generator().manual_seed(1)
    mu = torch.rand(n, 3, 24, 24, generator=g).requires_grad_(True)
    hr, lr = torch.rand(n, 3, 24, 24, generator=g), torch.rand(n, 3, 12, 12, generator=g)
    kinfo = torch.tensor([[1.0, 2.0, 0.999], [2.0, 1.0, -0.999]], requires_grad=True)
    kgt = torch.tensor([[1.5, 1.5, 0.2], [1.0, 2.0, -0.1]])
    sig = torch.full((n, 1, 1, 1), 1e-3, requires_grad=True)
    rho_draw = torch.tensor([[3.0], [-3.0]])                                      # 0.999 + 0.03 > 1, -0.999 - 0.03 < -1
    loss, _ = O.elbo_sisr(mu, sig, kinfo, hr, lr, sig.detach(), 40.5, kgt, 50.0, 1e-4, 1e-5, sf, 21, [0.02, 2], False,
                          "Bicubic", gamma_draw=torch.full((n, 2), 49.0), rho_draw=rho_draw,
                          z_draw=torch.zeros(n, 3, 24, 24))
    loss.backward()
    expect = (kinfo.detach()[:, 2] - kgt[:, 2]) / 1e-4 * 0.02 * 2 / 3 / n
    torch.testing.assert_close(kinfo.grad[:, 2], expect, rtol=1e-5, atol=0)


def test_synth_sigma_map_spans_down_to_up_and_peaks_at_the_centre():
    p = 48
    patch = np.zeros((p, p, 3), dtype=np.uint8)
    params = [10.3, 40.7, 20.0, 0.3, 0.05, 0.0]
    _, _, sg = O.synth_denoise_sample(patch, params, 0, np.zeros((p, p, 3), np.float32))
    s = sg[0].sqrt()
    assert abs(s.max().item() - 0.3) < 1e-6 and abs(s.min().item() - 0.05) < 1e-6
    iy, ix = divmod(int(s.argmax()), p)
    assert (iy, ix) == (10, 41)
    assert math.isclose(float(s[p - 1, 0]), 0.05, rel_tol=0, abs_tol=1e-6)        # the farthest corner
