"""Per-kernel parity on the B200, through the C ABI (ctypes): every implicit-GEMM geometry and
epilogue of vk_conv_igemm / vk_conv_wgrad against torch fp32 on operands pre-rounded to the MMA
operand format, plus the HBM-bound kernels against their closed forms.

Tolerances (written where they are applied, tools/debug_conv.py and tools/debug_wgrad.py):
  conv fprop/dgrad   rel-L2 < 2e-4 (tf32 operands, fp32 accumulate) / 5e-3 (bf16 OUTPUT rounding)
  wgrad              rel-L2 < 1e-4 (fp32 accumulate and fp32 output in both modes)
"""
import math
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))

pytestmark = pytest.mark.gpu

import debug_conv  # noqa: E402
import debug_wgrad  # noqa: E402

CONV_CASES = [(i, c["name"]) for i, c in enumerate(debug_conv.CASES) if not c.get("bench")]
WGRAD_CASES = [(i, c["name"]) for i, c in enumerate(debug_wgrad.CASES) if not c.get("bench")]


@pytest.fixture(scope="module", autouse=True)
def _needs_cuda():
    assert torch.cuda.is_available(), "gpu-marked tests need a CUDA device"
    from virnet_b200 import lib
    lib.load()


@pytest.mark.parametrize("idx,name", CONV_CASES, ids=[n for _, n in CONV_CASES])
def test_conv_igemm_case(idx, name):
    assert debug_conv.run_case(idx) == 0


@pytest.mark.parametrize("idx,name", WGRAD_CASES, ids=[n for _, n in WGRAD_CASES])
def test_conv_wgrad_case(idx, name):
    assert debug_wgrad.run_case(idx) == 0


def test_conv_full_size_linearity():
    """Size-independent property at the bench shape (16x96x128x128): conv(a*x1 + x2) == a*conv(x1) + conv(x2)
    up to output rounding; checked in tf32 storage (fp32 outputs) so the property is tight."""
    from virnet_b200 import ops
    dt = ops.VK_TF32
    g = torch.Generator(device="cuda").manual_seed(5)
    q = lambda t: (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
    # power-of-two scale and operands with few mantissa bits keep a*x1 + x2 exactly representable in tf32
    x1 = (torch.randint(-8, 9, (4, 128, 128, 96), device="cuda", generator=g).float() / 8)
    x2 = (torch.randint(-8, 9, (4, 128, 128, 96), device="cuda", generator=g).float() / 8)
    w = q(torch.randn(96, 96, 3, 3, device="cuda", generator=g) / 30)
    wp = ops.pack_conv_weight(w, dt, 96)
    outs = []
    for x in (x1, x2, 2.0 * x1 + x2):
        o = torch.empty(4, 128, 128, 96, device="cuda")
        ops.conv_igemm(x.contiguous(), wp, dtype=dt, kind=ops.VK_CONV3X3_S1, cout=96, ldo=96, out1=o)
        outs.append(o)
    torch.cuda.synchronize()
    lin = 2.0 * outs[0] + outs[1]
    rel = ((outs[2] - lin).norm() / lin.norm()).item()
    assert rel < 1e-5, rel


def test_elbo_kernel_matches_golden(kat, golden_dir):
    """vk_elbo_denoise against the reference-generated fixture (tools/gen_golden.py): tolerance 1e-5 rel (fp32)."""
    from virnet_b200 import ops
    g = torch.Generator().manual_seed(7)
    mu = torch.rand(2, 3, 16, 16, generator=g)
    sg = torch.rand(2, 1, 16, 16, generator=g) * 0.05 + 1e-3
    y = torch.rand(2, 3, 16, 16, generator=g)
    gt = torch.rand(2, 3, 16, 16, generator=g)
    b0 = 24.5 * (torch.rand(2, 1, 16, 16, generator=g) * 0.05 + 1e-3)
    d_mu, d_sg = torch.empty_like(mu).cuda(), torch.empty_like(sg).cuda()
    dig = float(torch.digamma(torch.tensor(23.5, dtype=torch.float64)))
    out4 = ops.elbo_denoise(mu.cuda(), sg.cuda(), y.cuda(), gt.cuda(), b0.cuda(), eps2=1e-6, alpha0=24.5,
                            digamma_am1=dig, d_mu=d_mu, d_sigma=d_sg).cpu()
    k = kat["elbo_16"]
    for got, key in zip(out4.tolist(), ("loss", "lh", "kl_gauss", "kl_igamma")):
        assert abs(got - k[key]) <= 1e-5 * max(1.0, abs(k[key])), (key, got, k[key])
    ref = torch.load(golden_dir / "elbo_16.pt")
    torch.testing.assert_close(d_mu.cpu(), ref["d_mu"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(d_sg.cpu(), ref["d_sigma"], rtol=1e-5, atol=1e-6)


def test_elbo_kernel_3ch_sigma_and_beta0_scale():
    """sigma_chn = 3 (denoising-real layout) and the fused beta0 = alpha0 * sigma_gt scale."""
    from oracle import virnet_oracle as O
    from virnet_b200 import ops
    g = torch.Generator().manual_seed(11)
    mu = torch.rand(3, 3, 19, 23, generator=g).requires_grad_(True)
    sg = (torch.rand(3, 3, 19, 23, generator=g) * 0.05 + 1e-3).requires_grad_(True)
    y, gt = torch.rand(3, 3, 19, 23, generator=g), torch.rand(3, 3, 19, 23, generator=g)
    sgt = torch.rand(3, 3, 19, 23, generator=g) * 0.05 + 1e-3
    loss, lh, kg, ig = O.elbo_denoising_simple(mu, sg, y, gt, 1e-6, 24.5, 24.5 * sgt)
    loss.backward()
    d_mu, d_sg = torch.empty_like(mu).cuda(), torch.empty_like(sg).cuda()
    dig = float(torch.digamma(torch.tensor(23.5, dtype=torch.float64)))
    out4 = ops.elbo_denoise(mu.detach().cuda(), sg.detach().cuda(), y.cuda(), gt.cuda(), sgt.cuda(), beta0_scale=24.5,
                            eps2=1e-6, alpha0=24.5, digamma_am1=dig, d_mu=d_mu, d_sigma=d_sg).cpu()
    want = torch.stack([loss, lh, kg, ig]).detach()
    torch.testing.assert_close(out4, want, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(d_mu.cpu(), mu.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(d_sg.cpu(), sg.grad, rtol=1e-4, atol=1e-6)


def test_adam_clip_step_matches_torch():
    """vk_adam_clip_step == clip_grad_norm_ per group + torch.optim.Adam (train_denoising_syn.py:182-184)."""
    import ctypes as C
    from virnet_b200 import lib, ops
    g = torch.Generator().manual_seed(3)
    n1, n2 = 1000, 50000
    p0 = torch.randn(n1 + n2, generator=g)
    params_t = [p0[:n1].clone().requires_grad_(True), p0[n1:].clone().requires_grad_(True)]
    opt = torch.optim.Adam(params_t, lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    flat_p = p0.clone().cuda()
    m, v = torch.zeros_like(flat_p), torch.zeros_like(flat_p)
    arr = (lib.vk_adam_group * 2)()
    arr[0].begin, arr[0].end, arr[0].max_norm = 0, n1, 1e2
    arr[1].begin, arr[1].end, arr[1].max_norm = n1, n1 + n2, 1e3
    groups = torch.frombuffer(bytearray(bytes(memoryview(arr))), dtype=torch.uint8).cuda()
    sq = ops.adam_ws(2, "cuda")
    norms = torch.zeros(2, device="cuda")
    for step in range(1, 4):
        # group 0 gets clipped (norm >> 1e2), group 1 does not (norm << 1e3); world-size-2 style sum + 0.5 scale
        gr = torch.cat([torch.randn(n1, generator=g) * 50, torch.randn(n2, generator=g) * 0.5])
        params_t[0].grad, params_t[1].grad = gr[:n1].clone(), gr[n1:].clone()
        tn0 = torch.nn.utils.clip_grad_norm_([params_t[0]], 1e2)
        tn1 = torch.nn.utils.clip_grad_norm_([params_t[1]], 1e3)
        opt.step()
        ops.adam_clip_step(flat_p, (2.0 * gr).cuda(), m, v, groups, 2, n2, sq, grad_scale=0.5, lr=1e-3, beta1=0.9,
                           beta2=0.999, eps=1e-8, step=step, norms_out=norms)
        torch.testing.assert_close(norms.cpu(), torch.stack([tn0, tn1]), rtol=1e-5, atol=0)
        want = torch.cat([params_t[0].detach(), params_t[1].detach()])
        torch.testing.assert_close(flat_p.cpu(), want, rtol=1e-5, atol=1e-6)


def test_pack_input_reflect_sqrt_concat():
    """vk_pack_input == cat[pad_input(x), pad_input(sqrt(sigma))] in NHWC (AttResUNet.py:147-153, util_net.py:20-25)."""
    from oracle import virnet_oracle as O
    from virnet_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, 37, 50, generator=g)
    s = torch.rand(2, 1, 37, 50, generator=g) + 0.1
    want = torch.cat([O.pad_input(x, 4), O.pad_input(s.sqrt(), 4)], 1)        # [2,4,40,52]
    for dt, tol in ((ops.VK_TF32, 1e-6), (ops.VK_BF16, 8e-3)):
        out = torch.full((2, 40, 52, ops.chan_pad(4, dt)), float("nan"), device="cuda", dtype=ops.TORCH_DTYPE[dt])
        ops.pack_input(x.cuda(), out, dtype=dt, extra=s.cuda(), extra_is_map=True, extra_sqrt_mask=1)
        got = out.float().cpu()
        assert torch.isfinite(got).all()
        torch.testing.assert_close(got[..., :4].permute(0, 3, 1, 2), want, rtol=tol, atol=tol)
        assert (got[..., 4:] == 0).all()


def test_elbo_list_form_of_mu_matches_the_reference_formula():
    """Deep-supervision list form (loss/ELBO_simple.py:30-36,44-49): lh and kl_gauss averaged over the list, kl_Igamma
    shared; value and gradients against the oracle's per-element restatement (fp32, 1e-5)."""
    from oracle import virnet_oracle as O
    from virnet_b200.loss.ELBO_simple import elbo_denoising_simple
    g = torch.Generator().manual_seed(3)
    mus = [torch.rand(2, 3, 16, 16, generator=g).requires_grad_(True) for _ in range(3)]
    sg = (torch.rand(2, 1, 16, 16, generator=g) * 0.05 + 1e-3).requires_grad_(True)
    y, gt = torch.rand(2, 3, 16, 16, generator=g), torch.rand(2, 3, 16, 16, generator=g)
    b0 = 24.5 * (torch.rand(2, 1, 16, 16, generator=g) * 0.05 + 1e-3)
    parts = [O.elbo_denoising_simple(m, sg, y, gt, 1e-6, 24.5, b0) for m in mus]
    lh = sum(p[1] for p in parts) / 3
    kg = sum(p[2] for p in parts) / 3
    want = lh + kg + parts[0][3]
    want.backward()
    mus_c = [m.detach().cuda().requires_grad_(True) for m in mus]
    sg_c = sg.detach().cuda().requires_grad_(True)
    loss, lh_c, kg_c, ig_c = elbo_denoising_simple(mus_c, sg_c, y.cuda(), gt.cuda(), 1e-6, 24.5, b0.cuda())
    loss.backward()
    assert abs(loss.item() - want.item()) <= 1e-5 * abs(want.item())
    assert abs(lh_c.item() - lh.item()) <= 1e-5 * abs(lh.item()) and abs(ig_c.item() - parts[0][3].item()) <= 1e-5 * abs(parts[0][3].item())
    for a, b in zip(mus_c, mus):
        torch.testing.assert_close(a.grad.cpu(), b.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(sg_c.grad.cpu(), sg.grad, rtol=1e-5, atol=1e-5)
