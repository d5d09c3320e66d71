#!/usr/bin/env python
"""Weight-gradient launches of the narrow layers (head 4->96, tail 96->3) at b=32, 128x128, bf16: the merged-tap and the
swapped-operand paths against the plain layout (VK_WGRAD_NO_MERGE=1 disables merging)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from virnet_b200 import ops  # noqa: E402
from tools.conv_bench import time_it  # noqa: E402

n, h = 32, 128
dt = ops.VK_BF16
big = torch.randn(n, h, h, 96, device="cuda").bfloat16()
small = torch.zeros(n, h, h, 16, device="cuda").bfloat16()
small[..., :4] = torch.randn(n, h, h, 4, device="cuda").bfloat16()
db96, db3 = torch.zeros(96, device="cuda"), torch.zeros(3, device="cuda")
# head: dY = big (Cout 96), X = small (Cin 4)
dw = torch.zeros(9, 96, 4, device="cuda")
us = time_it(lambda: ops.conv_wgrad(big, small, dw, dtype=dt, kind=ops.VK_CONV3X3_S1, m_valid=96, n_valid=4, dbias=db96))
print(f"head 4->96  (narrow N operand)        : {us:7.1f} us", flush=True)
# tail: dY = small (Cout 3), X = big (Cin 96)
dw = torch.zeros(9, 3, 96, device="cuda")
us = time_it(lambda: ops.conv_wgrad(small, big, dw, dtype=dt, kind=ops.VK_CONV3X3_S1, m_valid=3, n_valid=96, dbias=db3))
print(f"tail 96->3  plain (narrow M operand)  : {us:7.1f} us", flush=True)
us = time_it(lambda: ops.conv_wgrad(big, small, dw, dtype=dt, kind=ops.VK_CONV3X3_S1, m_valid=96, n_valid=3, swapped=True))
print(f"tail 96->3  swapped                   : {us:7.1f} us", flush=True)
ws = ops.channel_sum_ws(16, "cuda")
us = time_it(lambda: ops.channel_sum(small, 3, db3, dtype=dt, ws=ws))
print(f"channel_sum of the 16-channel dY      : {us:7.1f} us", flush=True)
