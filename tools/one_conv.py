#!/usr/bin/env python
"""One conv layer, a few launches — the command to wrap in `ncu --set full --import-source on`."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from virnet_b200 import ops  # noqa: E402
from tools.conv_bench import make  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--c", type=int, default=96)
ap.add_argument("--h", type=int, default=128)
ap.add_argument("--n", type=int, default=16)
ap.add_argument("--epi", default="resid_dual")
ap.add_argument("--impl", type=int, default=0)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--cout", type=int, default=0)
ap.add_argument("--s2", action="store_true", help="3x3 stride-2 kind")
args = ap.parse_args()
x, w, kw, o1, o2 = make(args.c, args.cout or args.c, args.h, args.n, args.epi, ops.VK_BF16,
                        ops.VK_CONV3X3_S2 if args.s2 else ops.VK_CONV3X3_S1)
for _ in range(args.iters):
    ops.conv_igemm(x, w, tune=dict(impl=args.impl), **kw)
torch.cuda.synchronize()
print("done")
