#!/usr/bin/env python
"""CUDA-graph training step under data parallelism (NCCL all-reduce captured inside the graph): run with
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/graph_ddp_check.py [batch_per_gpu]
Checks that all ranks hold identical parameters after graph-replayed steps and prints eager vs graph step times."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import virnet_b200  # noqa: E402
from virnet_b200.trainer import DenoiseTrainer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(1234)
net = virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=bench.N_FEAT, dep_S=bench.DEP_S, n_resblocks=bench.N_RES,
                                noise_cond=True, extra_mode="Input", noise_avg=False, precision="bf16").to(dev)
tr = DenoiseTrainer(net)
batch = bench.synth_batch(b, rank, dev)
res = {}
for name, fn in (("eager", tr.step), ("graph", tr.step_graph)):
    for _ in range(5):
        fn(*batch)
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(30):
        losses = fn(*batch)
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / 30], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[name] = t.item()
flat = tr.engine.flat_params
ref = flat.clone()
dist.broadcast(ref, 0)
same = torch.tensor([float(torch.equal(ref, flat))], device=dev)
dist.all_reduce(same, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps(dict(n_gpus=world, batch_per_gpu=b, eager_ms=round(res["eager"], 3), graph_ms=round(res["graph"], 3),
                          eager_patches_s=round(world * b / res["eager"] * 1e3, 1),
                          graph_patches_s=round(world * b / res["graph"] * 1e3, 1),
                          params_identical_across_ranks=bool(same.item()), finite=bool(torch.isfinite(flat).all().item()),
                          loss=losses[0].item())), flush=True)
dist.destroy_process_group()
