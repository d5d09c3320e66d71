"""Evaluation-side fusions on the B200 against the oracle's restatement of the reference functions
(oracle/eval_protocol.py, itself pinned on the reference by tests/test_oracle_eval.py):

* img_as_ubyte / calculate_psnr (bit-exact) / calculate_ssim (1e-9) on the reference's own test images, RGB and Y
  channel with border (utils/util_image.py:16-153);
* the 8-fold flip / rotate self-ensemble as one batched forward (scripts/denoising_virnet_real_sidd.py:120-136), square
  and non-square images;
* forward_chop quadrant tiling (utils/util_net.py:27-65) as one batched forward."""
import numpy as np
import pytest
import torch

from eval_common import cbsd68_images, kat, set5_images

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def K():
    return kat()


def test_device_metrics_equal_the_reference_functions(K):
    from oracle import eval_protocol as E
    from virnet_b200.utils import util_image as U
    rng = np.random.default_rng(0)
    pairs = []
    for im in cbsd68_images(K)[:3] + set5_images(K)[2:4]:
        noisy = np.clip(im.astype(np.float32) / 255.0 + rng.standard_normal(im.shape).astype(np.float32) * 0.1, 0, 1)
        pairs.append((noisy, im))
    for noisy, gt in pairs:
        x = torch.from_numpy(noisy.transpose(2, 0, 1)[None]).cuda()
        n8 = U.img_as_ubyte(x)[0]
        assert np.array_equal(n8.cpu().numpy(), E.img_as_ubyte(noisy))           # rint(x * 255) in fp32, like skimage
        g8 = torch.from_numpy(gt).cuda()
        for border, ycbcr in ((0, False), (4, True), (16, True), (7, False)):
            assert U.calculate_psnr(n8, g8, border, ycbcr) == E.calculate_psnr(E.img_as_ubyte(noisy), gt, border, ycbcr)
        for border, ycbcr in ((0, False), (16, True)):
            a, b = U.calculate_ssim(n8, g8, border, ycbcr), E.calculate_ssim(E.img_as_ubyte(noisy), gt, border, ycbcr)
            assert abs(a - b) < 1e-9, (a, b)
    # reference known answers (tests/golden/eval_kat.json) through the device path
    noisy = E.niid_noisy_images(cbsd68_images(K)[:2])
    for x, gt, e in zip(noisy, cbsd68_images(K), K["denoise"]):
        n8 = U.img_as_ubyte(torch.from_numpy(np.clip(x, 0, 1).transpose(2, 0, 1)[None]).cuda())[0]
        g8 = torch.from_numpy(gt).cuda()
        assert U.calculate_psnr(n8, g8, 0, False) == e["metric_kat"]["psnr_rgb"]
        assert U.calculate_psnr(n8, g8, 4, True) == e["metric_kat"]["psnr_y_b4"]
        assert abs(U.calculate_ssim(n8, g8, 16, True) - e["metric_kat"]["ssim_y_b16"]) < 1e-9
    assert U.calculate_psnr(g8, g8) == float("inf")
    # batch helpers (utils/util_image.py:91-116)
    xb = torch.rand(2, 3, 40, 52).cuda()
    yb = (xb + 0.05 * torch.randn_like(xb)).clamp(0, 1)
    want = np.mean([E.calculate_psnr(E.img_as_ubyte(xb[i].cpu().numpy().transpose(1, 2, 0)),
                                     E.img_as_ubyte(yb[i].cpu().numpy().transpose(1, 2, 0))) for i in range(2)])
    assert abs(U.batch_PSNR(yb, xb) - want) < 1e-12


@pytest.mark.parametrize("shape", [(2, 3, 32, 32), (1, 3, 28, 44)])
def test_self_ensemble_is_one_batched_forward_and_matches_the_reference_loop(shape):
    from oracle import eval_protocol as E
    from virnet_b200 import lib
    from virnet_b200.utils.util_image import self_ensemble
    import virnet_b200
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=[32, 64, 96], dep_S=5, n_resblocks=2, noise_cond=True,
                                    extra_mode="Input", noise_avg=False, precision="tf32").cuda().eval()
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(4)).cuda()
    l0 = lib.launch_count()
    out = self_ensemble(net, x)
    launches = lib.launch_count() - l0
    with torch.no_grad():
        l1 = lib.launch_count()
        net(x)
        one_forward = lib.launch_count() - l1
    assert launches <= (1 if shape[2] == shape[3] else 2) * one_forward + 4, (launches, one_forward)   # not 8 forwards
    # the reference loop: eight separate forwards, numpy flips, float32 accumulation, / 8

    def fn(im_hwc):
        t = torch.from_numpy(np.ascontiguousarray(im_hwc.transpose(2, 0, 1))[None]).cuda()
        with torch.no_grad():
            return net(t)[0][0].cpu().numpy().transpose(1, 2, 0)

    for i in range(shape[0]):
        want = E.self_ensemble(fn, x[i].cpu().numpy().transpose(1, 2, 0))
        got = out[i].cpu().numpy().transpose(1, 2, 0)
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)


def test_aug8_kernel_is_data_aug_np():
    from oracle import virnet_oracle as O
    from virnet_b200 import lib
    from virnet_b200.ops import _ptr, _stream
    x = torch.rand(2, 3, 5, 7).cuda()
    a = torch.empty(4, 2, 3, 5, 7).cuda()
    b = torch.empty(4, 2, 3, 7, 5).cuda()
    lib.check(lib.load().vk_aug8(_ptr(x), _ptr(a), _ptr(b), 6, 5, 7, _stream()), "vk_aug8")
    for n in range(2):
        hwc = x[n].cpu().numpy().transpose(1, 2, 0)
        for slot, mode in enumerate((0, 1, 4, 5)):
            assert np.array_equal(a[slot, n].cpu().numpy().transpose(1, 2, 0), O.data_aug_np(hwc, mode)), mode
        for slot, mode in enumerate((2, 3, 6, 7)):
            assert np.array_equal(b[slot, n].cpu().numpy().transpose(1, 2, 0), O.data_aug_np(hwc, mode)), mode


def test_forward_chop_matches_the_reference_tiling():
    from oracle import eval_protocol as E
    from virnet_b200.utils.util_net import forward_chop
    import virnet_b200
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNetSR(im_chn=3, n_feat=[32, 64, 96], dep_K=2, n_resblocks=1, extra_mode="Both",
                                      precision="tf32").cuda().eval()
    x = torch.rand(1, 3, 50, 66, generator=torch.Generator().manual_seed(2)).cuda()
    with torch.no_grad():
        got = forward_chop(lambda t: net(t, 2)[0], x, scale=2, shave=6)
        want = E.forward_chop(lambda t: net(t.contiguous().cuda(), 2)[0].cpu(), x.cpu(), scale=2, shave=6)
    assert got.shape == (1, 3, 100, 132)
    assert rel(got.cpu(), want) < 1e-6
    # and the recursive branch (which the reference cannot execute: it drops the `net` argument) tiles 16 patches
    with torch.no_grad():
        rec = forward_chop(lambda t: net(t, 2)[0], x, scale=2, shave=6, min_size=600)
    assert rec.shape == got.shape and torch.isfinite(rec).all()
