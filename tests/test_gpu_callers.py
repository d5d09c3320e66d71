"""Caller-level drop-in proof: the reference's OWN caller code runs unchanged on top of
virnet_b200 when only the import of networks.VIRNet / loss.ELBO_simple is swapped.

* scripts/testing_demo.py: `load_model` (constructor call, `.cuda()`, checkpoint dict with and without DDP's
  `module.` prefix, `load_state_dict(strict=True)`, `.eval()`) and `process_image` (numpy in, `no_grad`, in-place
  `clamp_` on the returned tensor, numpy out) are executed from the reference file itself;
* train_denoising_syn.py:169-184: the loop body (data to GPU, beta0, zero_grad, forward, elbo_denoising_simple,
  backward, two clip_grad_norm_, optimizer.step) is exec'd VERBATIM from the reference source file, once with the
  reference's modules and once with ours, from the same weights and batch; the updated parameters must agree;
* torch DDP: `DDP(net, device_ids=[rank])` (train_denoising_syn.py:70-71) wraps the module; gradients through the
  DDP hooks equal the un-wrapped ones (single-process group here; tools/ddp_equiv_check.py is the 2-GPU version).

Needs the reference sources (/root/reference, or baseline/_ref staged by build()); skipped when neither exists.
The reference side runs on the same GPU in true fp32 (TF32 off) as the yardstick; ours in tf32 mode (1e-3 bar)."""
import importlib.util
import os
import sys
import textwrap
import types
from collections import OrderedDict
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
import ref_import  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_import.available(), reason="reference sources not staged (baseline/_ref)")]


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture()
def fp32_reference():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _swapped_modules(monkeypatch, precision="tf32"):
    """What a maintainer's one-line import change amounts to: `networks.VIRNet` and
    `loss.ELBO_simple` resolve to virnet_b200's modules."""
    import virnet_b200
    import virnet_b200.loss.ELBO_simple as our_loss
    monkeypatch.setenv("VIRNET_B200_PRECISION", precision)
    shim = types.ModuleType("networks.VIRNet")
    shim.VIRAttResUNet, shim.VIRAttResUNetSR = virnet_b200.VIRAttResUNet, virnet_b200.VIRAttResUNetSR
    monkeypatch.setitem(sys.modules, "networks.VIRNet", shim)
    monkeypatch.setitem(sys.modules, "loss.ELBO_simple", our_loss)


def _load_testing_demo():
    ref_import.import_reference()                      # stubs thop / lpips / skimage, puts the reference on sys.path
    spec = importlib.util.spec_from_file_location("ref_testing_demo", ref_import.REF_ROOT / "scripts" / "testing_demo.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("task,sf,shape,prefix", [("denoising-syn", None, (45, 62), ""),
                                                  ("denoising-real", None, (40, 56), "module."),
                                                  ("sisr", 4, (23, 30), "")])
def test_reference_testing_demo_runs_on_the_drop_in(task, sf, shape, prefix, tmp_path, monkeypatch, fp32_reference):
    demo = _load_testing_demo()
    vir, _ = ref_import.import_reference()
    # a checkpoint in the reference's format, from the reference's own module (seed-1234 init; model_zoo is empty)
    torch.manual_seed(1234)
    if task == "denoising-syn":
        ref = vir.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=[96, 192, 288], dep_S=5, n_resblocks=3, noise_cond=True,
                                extra_mode="Input", noise_avg=False)
    elif task == "denoising-real":
        ref = vir.VIRAttResUNet(im_chn=3, sigma_chn=3, n_feat=[96, 160, 224, 288], dep_S=8, n_resblocks=3,
                                noise_cond=True, extra_mode="Input", noise_avg=False)
    else:
        ref = vir.VIRAttResUNetSR(im_chn=3, sigma_chn=1, dep_S=5, dep_K=8, n_feat=[96, 160, 224], n_resblocks=2,
                                  extra_mode="Both", noise_avg=True, noise_cond=True, kernel_cond=True)
    ckpt = tmp_path / "ckpt.pth"
    torch.save({"model_state_dict": OrderedDict((prefix + k, v) for k, v in ref.state_dict().items())}, ckpt)
    rng = np.random.default_rng(0)
    im_lq = rng.random((*shape, 3), dtype=np.float32)
    # reference end to end (its own modules, same GPU, fp32)
    out_ref = demo.process_image(demo.load_model(task, str(ckpt), sf), im_lq, task, sf)
    # the same caller code with the import swapped
    _swapped_modules(monkeypatch)
    net = demo.load_model(task, str(ckpt), sf)
    import virnet_b200
    assert isinstance(net, (virnet_b200.VIRAttResUNet, virnet_b200.VIRAttResUNetSR)) and not net.training
    out = demo.process_image(net, im_lq, task, sf)
    assert out.shape == out_ref.shape and out.dtype == out_ref.dtype
    assert 0.0 <= out.min() and out.max() <= 1.0
    assert rel(torch.from_numpy(out), torch.from_numpy(out_ref)) < 1e-3


def _train_loop_body():
    """train_denoising_syn.py:169-184, cut out of the reference file by content (first line `im_noisy, im_gt,
    sigma_gt = ...` to `optimizer.step()`), dedented — executed verbatim."""
    lines = (ref_import.REF_ROOT / "train_denoising_syn.py").read_text().splitlines()
    a = next(i for i, l in enumerate(lines) if "im_noisy, im_gt, sigma_gt = [x.cuda(rank) for x in data]" in l)
    b = next(i for i in range(a, len(lines)) if lines[i].strip() == "optimizer.step()")
    body = textwrap.dedent("\n".join(lines[a:b + 1]))
    assert "elbo_denoising_simple(" in body and "clip_grad_norm_" in body and "loss.backward()" in body
    return body


def _run_reference_loop(net, loss_fn, data, steps=2):
    from torch import nn
    import torch.optim as optim
    args = {"eps2": 1e-6, "clip_grad_R": 1e3, "clip_grad_S": 1e2, "var_window": 7, "lr": 1e-4}
    optimizer = optim.Adam(net.parameters(), lr=args["lr"])                        # train_denoising_syn.py:74
    param_R = [x for name, x in net.named_parameters() if "rnet" in name.lower()]   # :153-154
    param_S = [x for name, x in net.named_parameters() if "snet" in name.lower()]
    alpha0 = 0.5 * torch.tensor([args["var_window"] ** 2], dtype=torch.float32).cuda()   # :157
    ns = dict(net=net, data=data, rank=0, alpha0=alpha0, args=args, optimizer=optimizer, nn=nn, param_R=param_R,
              param_S=param_S, elbo_denoising_simple=loss_fn, torch=torch)
    body = _train_loop_body()
    net.train()
    losses = []
    for _ in range(steps):
        exec(body, ns)                                                               # noqa: S102
        losses.append([float(ns[k]) for k in ("loss", "g_lh", "kl_g", "kl_Igam")])
        norms = (float(ns["total_norm_R"]), float(ns["total_norm_S"]))
    return losses, norms


def test_reference_training_loop_body_runs_verbatim_on_the_drop_in(monkeypatch, fp32_reference):
    vir, elbo = ref_import.import_reference()
    import virnet_b200
    from virnet_b200.loss.ELBO_simple import elbo_denoising_simple as our_loss
    kw = dict(im_chn=3, sigma_chn=1, n_feat=[32, 64, 96], dep_S=5, n_resblocks=2, noise_cond=True, extra_mode="Input",
              noise_avg=False)
    torch.manual_seed(1234)
    ref = vir.VIRAttResUNet(**kw).cuda()
    ours = virnet_b200.VIRAttResUNet(**kw, precision="tf32")
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours = ours.cuda()
    g = torch.Generator().manual_seed(0)
    im_gt = torch.rand(4, 3, 64, 64, generator=g)
    sig = 5 / 255 + torch.rand(4, 1, 64, 64, generator=g) * 70 / 255
    data = [im_gt + torch.randn(4, 3, 64, 64, generator=g) * sig, im_gt, (sig ** 2).clamp_min(1e-10)]
    l_ref, n_ref = _run_reference_loop(ref, elbo.elbo_denoising_simple, data)
    l_our, n_our = _run_reference_loop(ours, our_loss, data)
    for a, b in zip(l_our, l_ref):
        for x, y in zip(a, b):
            assert abs(x - y) <= 2e-3 * abs(y) + 1e-6, (l_our, l_ref)
    for x, y in zip(n_our, n_ref):
        assert abs(x - y) <= 1e-2 * y, (n_our, n_ref)
    # after two optimizer steps the parameters moved the same way: compare the UPDATE (Adam's first steps have
    # magnitude lr per element, sign-driven, so compare where the reference gradient is not negligible)
    agree, total = 0, 0
    torch.manual_seed(1234)
    init = vir.VIRAttResUNet(**kw).state_dict()
    for (k, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
        du, dr = (p.detach().cpu() - init[k]), (q.detach().cpu() - init[k])
        big = dr.abs() > 1.5e-4                      # both steps pushed the same way in the reference
        agree += int(((du - dr).abs() < 5e-5)[big].sum())
        total += int(big.sum())
    assert total > 1000 and agree / total > 0.98, (agree, total)


def test_torch_ddp_wraps_the_drop_in_module(monkeypatch):
    """DDP(net, device_ids=[rank]) as in train_denoising_syn.py:70-71 (one-process NCCL group): forward through the
    DDP wrapper, backward through its reducer hooks, gradients identical to the un-wrapped module."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    import virnet_b200
    from virnet_b200.loss.ELBO_simple import elbo_denoising_simple
    if dist.is_initialized():
        pytest.skip("a process group already exists in this process")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29653")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        kw = dict(im_chn=3, sigma_chn=1, n_feat=[32, 64, 96], dep_S=5, n_resblocks=2, noise_cond=True,
                  extra_mode="Input", noise_avg=False, precision="tf32")
        torch.manual_seed(1234)
        net = virnet_b200.VIRAttResUNet(**kw).cuda()
        g = torch.Generator().manual_seed(0)
        im_gt = torch.rand(2, 3, 32, 32, generator=g)
        sig = 5 / 255 + torch.rand(2, 1, 32, 32, generator=g) * 70 / 255
        x = (im_gt + torch.randn(2, 3, 32, 32, generator=g) * sig).cuda()
        beta0 = (24.5 * (sig ** 2)).cuda()

        def grads(module):
            module.zero_grad(set_to_none=True)
            mu, sigma = module(x)
            elbo_denoising_simple(mu, sigma, x, im_gt.cuda(), 1e-6, 24.5, beta0)[0].backward()
            return {k.replace("module.", ""): p.grad.clone() for k, p in module.named_parameters()}

        plain = grads(net)
        wrapped = grads(DDP(net, device_ids=[0]))
        assert set(plain) == set(wrapped)
        for k in plain:
            assert torch.equal(plain[k], wrapped[k]) or rel(wrapped[k], plain[k]) < 1e-6, k
    finally:
        dist.destroy_process_group()
