"""Super-resolution forward path on the B200: the drop-in VIRAttResUNetSR
(KNet + SFT-modulated RNet through the C ABI) against the reference-generated fixture (kat.json /
sisr_x4_48_slices.pt, tools/gen_golden.py) and against the CPU oracle on the same weights.

Tolerances: 1e-3 relative in "tf32" mode (north_star's forward bar), 1e-2 in "bf16"."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(n_feat=(96, 160, 224), n_resblocks=2)


def make_sr(precision, n_feat=CFG["n_feat"], n_res=CFG["n_resblocks"], dep_K=8):
    import virnet_b200
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNetSR(im_chn=3, sigma_chn=1, dep_S=5, dep_K=dep_K, n_feat=list(n_feat),
                                      n_resblocks=n_res, extra_mode="Both", noise_avg=True, noise_cond=True,
                                      kernel_cond=True, precision=precision)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return net.cuda().eval(), sd


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("precision,tol", [("tf32", 1e-3), ("bf16", 1e-2)])
def test_sisr_x4_forward_vs_reference_fixture(precision, tol, kat, golden_dir):
    net, _ = make_sr(precision)
    x = torch.rand(2, 3, 48, 48, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        mu, kinfo, sigma = net(x.cuda(), 4)
    k = kat["sisr_x4_48"]
    assert mu.shape == (2, 3, 192, 192) and kinfo.shape == (2, 3) and sigma.shape == (2, 1, 1, 1)
    torch.testing.assert_close(kinfo.cpu(), torch.tensor(k["kinfo"]), rtol=tol * 5, atol=tol)
    torch.testing.assert_close(sigma.flatten().cpu(), torch.tensor(k["sigma"]), rtol=tol * 5, atol=tol)
    assert abs(mu.double().mean().item() - k["mu"]["mean"]) < tol * abs(k["mu"]["mean"]) * 5
    sl = torch.load(golden_dir / "sisr_x4_48_slices.pt")
    assert rel(mu[0, :, :8, :8].cpu(), sl["mu_slice"]) < tol * 5


@pytest.mark.parametrize("shape,sf", [((1, 3, 21, 30), 4), ((2, 3, 24, 24), 2)])
def test_sisr_forward_vs_oracle_ragged(shape, sf):
    """Odd LR sizes (reflect pad of the upsampled grid, crop) and another scale factor, small net."""
    from oracle import virnet_oracle as O
    net, sd = make_sr("tf32", n_feat=(32, 64, 96), n_res=2, dep_K=3)
    cfg = O.NetCfg(n_feat=(32, 64, 96), n_resblocks=2, extra_mode="Both", noise_avg=True, sisr=True, dep_K=3)
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        mu, kinfo, sigma = net(x.cuda(), sf)
        mu_o, kinfo_o, sigma_o = O.vir_sisr_forward(sd, x, sf, cfg)
    assert mu.shape == mu_o.shape
    assert rel(kinfo.cpu(), kinfo_o) < 1e-3 and rel(sigma.cpu(), sigma_o) < 1e-3
    assert rel(mu.cpu(), mu_o) < 1e-3


@pytest.mark.parametrize("precision,tol", [("tf32", 2e-2), ("bf16", 8e-2)])
@pytest.mark.parametrize("shape,sf,n_feat,dep_K", [((2, 3, 16, 16), 4, (32, 64, 96), 3), ((1, 3, 21, 30), 2, (32, 64, 96), 3),
                                                   ((2, 3, 16, 20), 4, (96, 160, 224), 8)])
def test_sisr_backward_vs_oracle(precision, tol, shape, sf, n_feat, dep_K):
    """Gradients of every parameter (SNet, KNet incl. the 9x9 head and channel attention, RNet incl. the SFT
    MLPs) of a random linear functional of (mu, kinfo, sigma) against torch autograd through the CPU oracle.
    Tolerance: relative L2 per sub-network, same bar as the denoising backward tests (tf32 2e-2, bf16 8e-2).
    The last case is the shipped sisr_x4.json width (96 / 160 / 224 features, 8 KNet blocks)."""
    from oracle import virnet_oracle as O
    net, sd = make_sr(precision, n_feat=n_feat, n_res=2, dep_K=dep_K)
    net.train()
    cfg = O.NetCfg(n_feat=n_feat, n_resblocks=2, extra_mode="Both", noise_avg=True, sisr=True, dep_K=dep_K)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(*shape, generator=g)
    N, C, h, w = shape
    w_mu = torch.randn(N, C, h * sf, w * sf, generator=g) / (h * sf)
    w_k = torch.randn(N, 3, generator=g)
    w_s = torch.randn(N, 1, 1, 1, generator=g) * 10

    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    mu_o, kinfo_o, sigma_o = O.vir_sisr_forward(sdr, x, sf, cfg)
    ((mu_o * w_mu).sum() + (kinfo_o * w_k).sum() + (sigma_o * w_s).sum()).backward()

    mu, kinfo, sigma = net(x.cuda(), sf)
    ((mu * w_mu.cuda()).sum() + (kinfo * w_k.cuda()).sum() + (sigma * w_s.cuda()).sum()).backward()
    got = {k: p.grad.detach().cpu() for k, p in net.named_parameters()}
    assert set(got) == set(sdr)
    for sub in ("SNet.", "KNet.", "RNet."):
        a = torch.cat([got[k].flatten() for k in sdr if k.startswith(sub)])
        b = torch.cat([sdr[k].grad.flatten() for k in sdr if k.startswith(sub)])
        assert rel(a, b) < tol, (sub, rel(a, b))
    # the small layers individually (their gradients are tiny next to the 3x3 convs' in the sub-network norm)
    groups = {"knet_head": ["KNet.head.weight"],
              "knet_ca": [k for k in sdr if k.startswith("KNet.body") and ".body.3." in k],
              "sft": [k for k in sdr if ".sft1." in k or ".sft2." in k]}
    for name, keys in groups.items():
        assert keys, name
        a = torch.cat([got[k].flatten() for k in keys])
        b = torch.cat([sdr[k].grad.flatten() for k in keys])
        assert rel(a, b) < tol, (name, rel(a, b))


def test_sisr_backward_then_inference_consistent():
    """A no-grad forward after a training forward/backward returns the same outputs (buffers are not clobbered)."""
    net, _ = make_sr("tf32", n_feat=(32, 64, 96), n_res=1, dep_K=2)
    net.train()
    x = torch.rand(1, 3, 16, 16, device="cuda")
    mu, kinfo, sigma = net(x, 4)
    mu_keep = mu.clone()
    (mu.sum() + kinfo.sum() + sigma.sum()).backward()
    with torch.no_grad():
        mu2, _, _ = net(x, 4)
    assert torch.equal(mu2, mu_keep)


def _sisr_batch(n, h, w, sf, seed=9):
    g = torch.Generator().manual_seed(seed)
    im_hr = torch.rand(n, 3, h * sf, w * sf, generator=g)
    im_lr = torch.nn.functional.avg_pool2d(im_hr, sf) + 0.02 * torch.randn(n, 3, h, w, generator=g)
    kinfo_gt = torch.stack([0.5 + 3 * torch.rand(n, generator=g), 0.5 + 3 * torch.rand(n, generator=g),
                            torch.rand(n, generator=g) - 0.5], dim=1)
    nlevel = (0.02 ** 2 * torch.ones(n, 1, 1, 1))
    return im_hr, im_lr, kinfo_gt, nlevel


def test_sisr_trainer_step_matches_oracle_step():
    """One SISRTrainer.step (forward, SISR ELBO, backward, per-sub-network clip, Adam; kernel-only) vs the oracle
    doing train_SISR.py:206-229 on the CPU with the same random draws; more steps must lower the loss."""
    from oracle import virnet_oracle as O
    from virnet_b200.trainer import SISRTrainer
    sf, n, h, w = 4, 2, 16, 16
    net, sd = make_sr("tf32", n_feat=(32, 64, 96), n_res=2, dep_K=3)
    net.train()
    cfg = O.NetCfg(n_feat=(32, 64, 96), n_resblocks=2, extra_mode="Both", noise_avg=True, sisr=True, dep_K=3)
    im_hr, im_lr, kinfo_gt, nlevel = _sisr_batch(n, h, w, sf)
    torch.manual_seed(77)
    draws = O.reference_draws(n, im_hr.shape, 50.0)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)
    mu, kinfo, sigma = O.vir_sisr_forward(params, im_lr, sf, cfg)
    alpha0 = 0.5 * 9 ** 2
    loss_o, det_o = O.elbo_sisr(mu, sigma, kinfo, im_hr, im_lr, nlevel, alpha0, kinfo_gt, 50.0, 1e-4, 1e-5, sf, 21,
                                [0.02, 2], False, "Bicubic", gamma_draw=draws[0], rho_draw=draws[1], z_draw=draws[2])
    loss_o.backward()
    norms_o = {nm: O.clip_grad_norm_([v for k, v in params.items() if k.lower().startswith(nm.lower())], mx).item()
               for nm, mx in (("RNet", 5e2), ("SNet", 1e2), ("KNet", 5e2))}
    opt.step()

    tr = SISRTrainer(net, sf, lr=1e-4, clip_grad_R=5e2, clip_grad_S=1e2, clip_grad_K=5e2)
    batch = [t.cuda() for t in (im_hr, im_lr, kinfo_gt, nlevel)]
    dd = tuple(d.cuda() for d in draws)
    terms = tr.step(*batch, draws=dd).clone()
    assert abs(terms[0].item() - loss_o.item()) < 5e-3 * abs(loss_o.item()), (terms.tolist(), loss_o.item())
    for i in range(1, 8):
        ref_v = det_o[i - 1].item()
        assert abs(terms[i].item() - ref_v) < 1e-2 * max(abs(ref_v), 1e-3), (i, terms[i].item(), ref_v)
    norms = dict(zip(tr.group_names, tr.grad_norms.tolist()))
    for nm in ("RNet", "SNet", "KNet"):
        assert abs(norms[nm] - norms_o[nm]) < 2e-2 * norms_o[nm], (nm, norms[nm], norms_o[nm])
    agree = tot = 0
    for name, p in net.named_parameters():
        d_ours = p.detach().cpu() - sd[name]
        d_ref = params[name].detach() - sd[name]
        big = d_ref.abs() > 0.5e-4
        agree += (torch.sign(d_ours[big]) == torch.sign(d_ref[big])).sum().item()
        tot += big.sum().item()
        assert (d_ours.abs() <= 1.0001e-4).all(), name
    assert agree / tot > 0.97, agree / tot
    first = terms[0].item()
    for _ in range(30):
        terms = tr.step(*batch, draws=dd)
    assert terms[0].item() < 0.9 * first, (first, terms[0].item())


def test_sisr_autograd_path_equals_trainer_path():
    """net(...) -> elbo_sisr -> loss.backward() (what train_SISR.py calls) fills p.grad with the gradients the
    fused trainer path uses."""
    from virnet_b200.loss.ELBO_simple import elbo_sisr, sisr_draws
    from virnet_b200.trainer import SISRTrainer
    sf, n, h, w = 2, 2, 16, 20
    net, _ = make_sr("tf32", n_feat=(32, 64, 96), n_res=1, dep_K=2)
    net.train()
    im_hr, im_lr, kinfo_gt, nlevel = [t.cuda() for t in _sisr_batch(n, h, w, sf)]
    torch.manual_seed(5)
    draws = sisr_draws(kinfo_gt, im_hr, 50.0)
    mu, kinfo, sigma = net(im_lr, sf)
    loss, detail = elbo_sisr(mu, sigma, kinfo, im_hr, im_lr, nlevel, torch.tensor([40.5]).cuda(), kinfo_gt,
                             torch.tensor([50.0]).cuda(), 1e-4, 1e-5, sf, 21, [0.02, 2], False, "Bicubic", draws=draws)
    loss.backward()
    g_autograd = torch.cat([p.grad.flatten() for p in net.parameters()])
    tr = SISRTrainer(net, sf, lr=0.0)
    eng = net.engine()
    terms = tr.step(im_hr, im_lr, kinfo_gt, nlevel, draws=draws)
    g_fused = torch.cat([eng.grad_view(p).flatten() for p in net.parameters()])
    assert abs(terms[0].item() - loss.item()) < 1e-5 * abs(loss.item())
    assert rel(g_fused, g_autograd) < 1e-5


def test_sisr_cuda_graph_step_matches_eager_step():
    """SISRTrainer.step_graph (one CUDA-graph launch; the loss's draws made outside the graph) follows the eager step."""
    from virnet_b200.loss.ELBO_simple import sisr_draws
    from virnet_b200.trainer import SISRTrainer
    sf, n, h, w = 4, 2, 16, 16
    batch = [t.cuda() for t in _sisr_batch(n, h, w, sf)]
    net_a, _ = make_sr("tf32", n_feat=(32, 64, 96), n_res=1, dep_K=2)
    net_b, _ = make_sr("tf32", n_feat=(32, 64, 96), n_res=1, dep_K=2)
    net_a.train(), net_b.train()
    tr_a, tr_b = SISRTrainer(net_a, sf, lr=1e-4), SISRTrainer(net_b, sf, lr=1e-4)
    torch.manual_seed(3)
    for it in range(4):
        draws = sisr_draws(batch[2], batch[0], 50.0)
        la = tr_a.step(*batch, lr=1e-4 * (1 + it), draws=draws).clone()
        lb = tr_b.step_graph(*batch, lr=1e-4 * (1 + it), draws=draws).clone()
        tol = 2e-4 if it == 0 else 2e-2
        torch.testing.assert_close(la, lb, rtol=tol, atol=1e-3)
    torch.manual_seed(4)
    first = tr_b.step_graph(*batch)[0].item()       # internal draws
    for _ in range(20):
        last = tr_b.step_graph(*batch)[0].item()
    assert last < first


# ------------------------------------------------------------------------------------------------------------------
# deterministic mode of the super-resolution step
# ------------------------------------------------------------------------------------------------------------------
def _run_sisr_det(mode, steps, n_feat, n_res, dep_K, precision, n, h, w, sf=4):
    from virnet_b200.loss.ELBO_simple import sisr_draws
    from virnet_b200.trainer import SISRTrainer
    net, _ = make_sr(precision, n_feat=n_feat, n_res=n_res, dep_K=dep_K)
    net.train()
    tr = SISRTrainer(net, sf, lr=1e-4, deterministic=True)
    batch = [t.cuda() for t in _sisr_batch(n, h, w, sf)]
    torch.manual_seed(11)
    hist = []
    for it in range(steps):
        draws = sisr_draws(batch[2], batch[0], 50.0)
        fn = tr.step_graph if mode == "graph" else tr.step
        terms = fn(*batch, lr=1e-4 * (1 + it), draws=draws)
        hist.append(torch.cat([terms.clone(), tr.grad_norms.clone()]))
    torch.cuda.synchronize()
    return torch.stack(hist).cpu(), torch.cat([p.detach().flatten() for p in net.parameters()]).cpu()


@pytest.mark.parametrize("n_feat,n_res,dep_K,precision,n,h,w", [((32, 64, 96), 1, 2, "tf32", 3, 16, 20),
                                                                ((96, 160, 224), 2, 8, "bf16", 4, 48, 48)])
def test_deterministic_sisr_training_is_bit_reproducible(n_feat, n_res, dep_K, precision, n, h, w):
    """SISRTrainer(deterministic=True): ordered split-K slabs for the convolutions, the fixed-order forms of the small
    per-sample kernels (SFT backward, AttLayer MLPs, CALayer, KNet head) and the atomic-free SISR loss make every loss
    term, gradient norm and parameter bit-identical from run to run, eager or CUDA-graph replayed.  The second case is
    the shipped x4 width (96/160/224, 8 KNet blocks) on 192x192 HR patches: 9 pixel chunks per sample in vk_sft_bwd_det."""
    h1, p1 = _run_sisr_det("eager", 3, n_feat, n_res, dep_K, precision, n, h, w)
    h2, p2 = _run_sisr_det("eager", 3, n_feat, n_res, dep_K, precision, n, h, w)
    h3, p3 = _run_sisr_det("graph", 3, n_feat, n_res, dep_K, precision, n, h, w)
    assert torch.isfinite(p1).all() and torch.isfinite(h1).all()
    assert torch.equal(h1, h2) and torch.equal(p1, p2), "eager runs differ"
    assert torch.equal(h1, h3) and torch.equal(p1, p3), "graph replay differs from the eager step"


def test_deterministic_sisr_gradients_match_the_atomic_path():
    """The fixed-order forms compute the same gradients as the atomic ones (to fp32 summation-order noise): every
    parameter of SNet / KNet / RNet and the loss terms, one step at lr = 0 on the same draws."""
    from virnet_b200.loss.ELBO_simple import sisr_draws
    from virnet_b200.trainer import SISRTrainer
    sf, n, h, w = 4, 3, 16, 20
    batch = [t.cuda() for t in _sisr_batch(n, h, w, sf)]
    torch.manual_seed(21)
    draws = sisr_draws(batch[2], batch[0], 50.0)
    out = []
    for det in (False, True):
        net, _ = make_sr("tf32", n_feat=(32, 64, 96), n_res=2, dep_K=3)
        net.train()
        tr = SISRTrainer(net, sf, deterministic=det)
        terms = tr.step(*batch, lr=0.0, draws=draws).clone()
        eng = tr.engine
        offs = eng.flat_offsets + [eng.flat_total]
        grads = {name: eng.flat_grads[offs[i]:offs[i + 1]].clone() for i, (name, _) in enumerate(net.named_parameters())}
        out.append((terms, grads, tr.grad_norms.clone()))
    (ta, ga, na), (tb, gb, nb) = out
    torch.testing.assert_close(ta, tb, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(na, nb, rtol=1e-4, atol=1e-6)
    for k in ga:
        assert rel(gb[k], ga[k]) < 1e-4, (k, rel(gb[k], ga[k]))
