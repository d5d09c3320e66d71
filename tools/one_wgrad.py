#!/usr/bin/env python
"""One wgrad layer, a few launches — the command to wrap in ncu."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from virnet_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--c", type=int, default=96)
ap.add_argument("--h", type=int, default=128)
ap.add_argument("--n", type=int, default=16)
args = ap.parse_args()
c, h, n = args.c, args.h, args.n
dy = torch.randn(n, h, h, c, device="cuda").bfloat16()
x = torch.randn(n, h, h, c, device="cuda").bfloat16()
dw = torch.zeros(9, c, c, device="cuda")
db = torch.zeros(c, device="cuda")
for _ in range(3):
    ops.conv_wgrad(dy, x, dw, dtype=ops.VK_BF16, kind=ops.VK_CONV3X3_S1, m_valid=c, n_valid=c, dbias=db)
torch.cuda.synchronize()
print("done")
