#!/usr/bin/env python
"""Which (kernel template arguments, epilogue combination) the persistent conv kernel is launched with in one training
step: run with VK_V2_DEBUG=1 and pipe stderr through this script's parser.  usage: VK_V2_DEBUG=1 python tools/v2_config_census.py [batch]"""
import re
import subprocess
import sys
from collections import Counter

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import torch
    import bench
    import virnet_b200
    from virnet_b200.trainer import DenoiseTrainer
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=bench.N_FEAT, dep_S=bench.DEP_S, n_resblocks=bench.N_RES,
                                    noise_cond=True, extra_mode="Input", noise_avg=False, precision=(sys.argv[3] if len(sys.argv) > 3 else "bf16")).to(dev)
    tr = DenoiseTrainer(net)
    batch = bench.synth_batch(int(sys.argv[2]), 0, dev)
    tr.step(*batch)
    torch.cuda.synchronize()
    print("CENSUS-START", file=sys.stderr, flush=True)
    tr.step(*batch)
    torch.cuda.synchronize()
    sys.exit(0)

b = sys.argv[1] if len(sys.argv) > 1 else "32"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
out = subprocess.run([sys.executable, __file__, "--child", b, prec], capture_output=True, text=True).stderr
out = out.split("CENSUS-START")[-1]
cnt = Counter()
for line in out.splitlines():
    m = re.search(r"kind=(\d+) .* (\d+)x(\d+) ldx=(\d+) wrows=(\d+) .* n_cta=(\d+) .* chunk=(\d+) nt=(\d+) .* pair=(\d) resident=(\d) fullk=(\d) mode=(-?\d+)", line)
    if m:
        kind, h, w, ldx, wrows, ncta, chunk, nt, pair, res, fullk, mode = m.groups()
        cnt[(f"chunk={chunk} nt={nt} pair={pair} fullk={fullk} mode={mode}", f"kind={kind} {h}x{w} {ldx}->{wrows} n_cta={ncta}")] += 1
for (k, shape), n in sorted(cnt.items()):
    print(f"{n:3d}  {k:48s} {shape}")
