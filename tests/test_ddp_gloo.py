"""World-size-2 data parallelism on CPU (gloo): the product's flat-bucket exchange
(virnet_b200/dp.py) must reproduce the reference's DDP semantics (train_denoising_syn.py:70-71,
133, 179-184): each rank takes global_batch // world patches, the loss is the LOCAL mean, gradients
are AVERAGED over ranks (== gradients of the global mean), and the per-sub-network clip runs on the
averaged gradients so every rank applies the identical update."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import virnet_oracle as O
from virnet_b200 import dp

CFG = O.NetCfg(n_feat=(16, 32, 48), n_resblocks=1, dep_S=3)
ALPHA0, EPS2 = 24.5, 1e-6


def _inputs(n):
    g = torch.Generator().manual_seed(7)
    im_gt = torch.rand(n, 3, 16, 16, generator=g)
    sig = 5 / 255 + torch.rand(n, 1, 16, 16, generator=g) * 70 / 255
    im_noisy = im_gt + torch.randn(n, 3, 16, 16, generator=g) * sig
    return im_noisy, im_gt, (sig ** 2).clamp_min(1e-10)


def _flat_grads(sd, sl):
    im_noisy, im_gt, sigma_gt = _inputs(4)
    _, _, _, grads = O.denoise_loss_and_grads(sd, CFG, im_noisy[sl], im_gt[sl], sigma_gt[sl], ALPHA0, EPS2)
    keys = list(sd.keys())
    return torch.cat([grads[k].reshape(-1) for k in keys]), keys


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    torch.manual_seed(1234)
    sd = O.build_state_dict(CFG)
    flat_p = torch.cat([v.reshape(-1) for v in sd.values()])
    if rank != 0:
        flat_p.add_(1.0)                       # diverge on purpose: the broadcast must restore rank 0's
    dp.broadcast_flat_params(flat_p, 0)
    b = dp.per_rank_batch(4, world)
    flat_g, _ = _flat_grads(sd, slice(rank * b, (rank + 1) * b))
    scale = dp.all_reduce_flat_grads(flat_g)
    # the gradient every rank must end up with: that of the mean over the GLOBAL batch, same weights
    ref_g, _ = _flat_grads(sd, slice(0, 4))
    out[rank] = (flat_p.clone(), flat_g * scale, ref_g)
    dist.destroy_process_group()


def test_flat_bucket_allreduce_matches_ddp_average():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    (p0, g0, ref_g), (p1, g1, _) = out[0], out[1]
    assert torch.equal(p0, p1)                 # broadcast restored rank 0's parameters on rank 1
    assert torch.equal(g0, g1)                 # every rank sees the same averaged gradient
    rel = ((g0 - ref_g).norm() / ref_g.norm()).item()
    assert rel < 1e-5, rel


def test_per_rank_batch_split():
    assert dp.per_rank_batch(16, 8) == 2 and dp.per_rank_batch(16, 1) == 16
    try:
        dp.per_rank_batch(4, 8)
    except ValueError:
        pass
    else:
        raise AssertionError("expected ValueError")


# ---- the same exchange under the SISR step (train_SISR.py:94-95, 206-229): three clip groups, per-rank loss draws ----
SR_CFG = O.NetCfg(n_feat=(16, 32, 48), n_resblocks=1, dep_S=3, extra_mode="Both", noise_avg=True, sisr=True, dep_K=2)


def _sr_flat_grads(sd, sl, draw_seed):
    g = torch.Generator().manual_seed(11)
    n, sf = 4, 2
    im_hr = torch.rand(n, 3, 24, 24, generator=g)
    im_lr = torch.nn.functional.avg_pool2d(im_hr, sf)
    kinfo_gt = torch.stack([1 + torch.rand(n, generator=g), 1 + torch.rand(n, generator=g), torch.rand(n, generator=g) - 0.5], 1)
    nlevel = torch.full((n, 1, 1, 1), 1e-3)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    mu, kinfo, sigma = O.vir_sisr_forward(params, im_lr[sl], sf, SR_CFG)
    m = mu.shape[0]
    dg = torch.Generator().manual_seed(draw_seed)           # every rank draws its own samples, like the reference
    draws = (torch._standard_gamma(torch.full((m, 2), 49.0), generator=dg), torch.randn(m, 1, generator=dg),
             torch.randn(mu.shape, generator=dg))
    loss, _ = O.elbo_sisr(mu, sigma, kinfo, im_hr[sl], im_lr[sl], nlevel[sl], 40.5, kinfo_gt[sl], 50.0, 1e-4, 1e-5, sf, 21,
                          [0.02, 2], False, "Bicubic", gamma_draw=draws[0], rho_draw=draws[1], z_draw=draws[2])
    loss.backward()
    return torch.cat([params[k].grad.reshape(-1) for k in sd]), loss.item()


def _sr_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    torch.manual_seed(1234)
    sd = O.build_state_dict(SR_CFG)
    b = dp.per_rank_batch(4, world)
    flat_g, _ = _sr_flat_grads(sd, slice(rank * b, (rank + 1) * b), draw_seed=100 + rank)
    local = flat_g.clone()
    scale = dp.all_reduce_flat_grads(flat_g)
    out[rank] = (local, flat_g * scale)
    dist.destroy_process_group()


def test_sisr_flat_bucket_allreduce_is_mean_of_rank_gradients():
    """With rank-local random draws the exchanged gradient must be exactly the mean of the two ranks' gradients (DDP's
    average), identical on both ranks, and non-zero for every sub-network (SNet / KNet / RNet all reach the loss)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sr_worker, args=(2, port, out), nprocs=2, join=True)
    (l0, g0), (l1, g1) = out[0], out[1]
    assert torch.equal(g0, g1)
    torch.testing.assert_close(g0, (l0 + l1) / 2, rtol=1e-6, atol=1e-12)
    torch.manual_seed(1234)
    sd = O.build_state_dict(SR_CFG)
    o = 0
    seen = {}
    for k, v in sd.items():
        sub = k.split(".")[0]
        seen[sub] = seen.get(sub, 0.0) + g0[o:o + v.numel()].abs().sum().item()
        o += v.numel()
    assert set(seen) == {"SNet", "KNet", "RNet"} and all(val > 0 for val in seen.values()), seen
