#!/usr/bin/env python
"""Multi-GPU checks of the data-parallel step (run under torchrun, one rank per GPU, NCCL):

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_equiv_check.py

1. torch DDP equivalence: gradients of `DDP(net, device_ids=[rank])` (train_denoising_syn.py:70-71; the reference's
   exchange) == the engine's flat gradient buffer after the bucketed all-reduce, times 1/world;
2. the overlapped, bucketed all-reduce (dp.BucketedGradSync) gives the same training trajectory as the single blocking
   all-reduce (VIRNET_B200_OVERLAP_ALLREDUCE=0 path), and all ranks hold identical parameters after the steps;
3. step_graph at world > 1 (graph = forward + ELBO + backward; NCCL + clip/Adam outside) follows the eager steps;
4. timings (CUDA events, max over ranks): b=32 eager with / without overlap, b=2 eager vs graph (the reference's
   per-GPU share of its global batch 16 on 8 GPUs).
Prints one JSON line per check on rank 0."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: E402

import bench  # noqa: E402
import virnet_b200  # noqa: E402
from virnet_b200 import dp  # noqa: E402
from virnet_b200.loss.ELBO_simple import elbo_denoising_simple  # noqa: E402
from virnet_b200.trainer import DenoiseTrainer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def say(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def make(prec="bf16", n_feat=bench.N_FEAT, n_res=bench.N_RES):
    torch.manual_seed(1234)
    return virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=n_feat, dep_S=5, n_resblocks=n_res, noise_cond=True,
                                     extra_mode="Input", noise_avg=False, precision=prec).to(dev)


def timed(fn, steps=20, warmup=5):
    for _ in range(warmup):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


# ---- 1. torch DDP == flat-bucket all-reduce ----
net = make("tf32", [32, 64, 96], 2)
batch = bench.synth_batch(8, rank, dev)
x, gt, sg = batch


def loss_of(module):
    mu, sigma = module(x)
    return elbo_denoising_simple(mu, sigma, x, gt, bench.EPS2, bench.ALPHA0, bench.ALPHA0 * sg)[0]


ddp = DDP(net, device_ids=[local])
ddp.zero_grad(set_to_none=True)
loss_of(ddp).backward()
g_ddp = torch.cat([p.grad.flatten() for p in net.parameters()])
tr = DenoiseTrainer(net, lr=0.0)
assert tr._sync is not None
tr._device_fwd_bwd(x, gt, sg, None, None, overlap=True)
torch.cuda.synchronize()
eng = tr.engine
g_ours = torch.cat([eng.grad_view(p).flatten() for p in net.parameters()]) / world
say(check="ddp_vs_flat_bucket", world=world, rel=rel(g_ours, g_ddp), buckets=[b[2:] for b in tr._sync.buckets_last_step])
assert rel(g_ours, g_ddp) < 1e-4
del ddp, tr, net

# ---- 2. overlapped buckets == single blocking all-reduce; ranks stay in lock step ----
params = {}
for mode in ("overlap", "blocking"):
    os.environ["VIRNET_B200_OVERLAP_ALLREDUCE"] = "1" if mode == "overlap" else "0"
    net = make("tf32", [32, 64, 96], 2)
    tr = DenoiseTrainer(net, lr=1e-3)
    assert (tr._sync is not None) == (mode == "overlap")
    for _ in range(5):
        tr.step(x, gt, sg)
    torch.cuda.synchronize()
    params[mode] = tr.engine.flat_params.clone()
    gathered = [torch.empty_like(params[mode]) for _ in range(world)]
    dist.all_gather(gathered, params[mode])
    same = all(torch.equal(gathered[0], g) for g in gathered)
    say(check=f"ranks_identical_{mode}", ok=bool(same))
    assert same
    del tr, net
os.environ["VIRNET_B200_OVERLAP_ALLREDUCE"] = "1"
# split-K weight gradients are accumulated with fp32 atomics (order-dependent in the last bits), Adam amplifies the sign
# of tiny gradients: compare the parameter UPDATE direction statistically
d = (params["overlap"] - params["blocking"]).abs()
say(check="overlap_vs_blocking", max_abs_diff=d.max().item(), frac_gt_1e_4=(d > 1e-4).float().mean().item())
assert (d > 1e-3).float().mean().item() < 1e-3

# ---- 3. step_graph at world > 1 ----
net_a, net_b = make("tf32", [32, 64, 96], 2), make("tf32", [32, 64, 96], 2)
tr_a, tr_b = DenoiseTrainer(net_a, lr=1e-3), DenoiseTrainer(net_b, lr=1e-3)
for it in range(4):
    la = tr_a.step(x, gt, sg).clone()
    lb = tr_b.step_graph(x, gt, sg).clone()
torch.cuda.synchronize()
d = (tr_a.engine.flat_params - tr_b.engine.flat_params).abs()
say(check="graph_vs_eager_world>1", loss_eager=la[0].item(), loss_graph=lb[0].item(),
    frac_gt_1e_3=(d > 1e-3).float().mean().item())
assert abs(la[0].item() - lb[0].item()) <= 2e-2 * abs(la[0].item()) and (d > 1e-3).float().mean().item() < 0.05
del tr_a, tr_b, net_a, net_b
torch.cuda.empty_cache()

# ---- 3b. deterministic mode at world > 1: run-to-run bit-identical, eager == graph, all ranks identical ----
def det_run(mode, steps=4):
    net = make("bf16", [32, 64, 96], 2)
    tr = DenoiseTrainer(net, lr=1e-3, deterministic=True)
    for it in range(steps):
        (tr.step if mode == "eager" else tr.step_graph)(x, gt, sg, lr=1e-3 * (1 + it))
    torch.cuda.synchronize()
    return tr.engine.flat_params.clone(), tr.losses.clone()


p1, l1 = det_run("eager")
p2, l2 = det_run("eager")
p3, l3 = det_run("graph")
gathered = [torch.empty_like(p1) for _ in range(world)]
dist.all_gather(gathered, p1)
say(check="deterministic_world>1", run_to_run_equal=bool(torch.equal(p1, p2) and torch.equal(l1, l2)),
    graph_equals_eager=bool(torch.equal(p1, p3)), max_abs_graph_vs_eager=(p1 - p3).abs().max().item(),
    ranks_identical=bool(all(torch.equal(gathered[0], g) for g in gathered)))
assert torch.equal(p1, p2) and torch.equal(l1, l2)
del p1, p2, p3
torch.cuda.empty_cache()

# ---- 4. timings, full-size network ----
for b in (32, 2):
    batch = bench.synth_batch(b, rank, dev)
    res = {"check": "timing", "world": world, "batch_per_gpu": b}
    for mode in ("overlap", "blocking"):
        os.environ["VIRNET_B200_OVERLAP_ALLREDUCE"] = "1" if mode == "overlap" else "0"
        tr = DenoiseTrainer(make("bf16"), lr=1e-4)
        ms = timed(lambda: tr.step(*batch))
        res[f"eager_{mode}_ms"] = round(ms, 4)
        res[f"eager_{mode}_patches_s"] = round(world * b / ms * 1e3, 1)
        if mode == "overlap":
            msg = timed(lambda: tr.step_graph(*batch))
            res["graph_ms"] = round(msg, 4)
            res["graph_patches_s"] = round(world * b / msg * 1e3, 1)
        del tr
        torch.cuda.empty_cache()
    os.environ["VIRNET_B200_OVERLAP_ALLREDUCE"] = "1"
    say(**res)
dist.destroy_process_group()
