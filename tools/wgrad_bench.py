#!/usr/bin/env python
"""Times vk_conv_wgrad at the bench layer shapes (tuning aid)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from virnet_b200 import ops  # noqa: E402
from tools.conv_bench import time_it  # noqa: E402


def run(c, h, n, tune=None, dt=ops.VK_BF16):
    tdt = ops.TORCH_DTYPE[dt]
    dy = torch.randn(n, h, h, c, device="cuda").to(tdt)
    x = torch.randn(n, h, h, c, device="cuda").to(tdt)
    dw = torch.zeros(9, c, c, device="cuda")
    db = torch.zeros(c, device="cuda")
    f = lambda: ops.conv_wgrad(dy, x, dw, dtype=dt, kind=ops.VK_CONV3X3_S1, m_valid=c, n_valid=c, dbias=db, tune=tune)
    us = time_it(f)
    flops = 2.0 * n * h * h * 9 * c * c
    print(f"wgrad C={c} {h}x{h} n={n} tune={tune}: {us:7.1f} us {flops / us / 1e6:6.0f} TF/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        # python tools/wgrad_bench.py --one N K_ROWS   (K_ROWS 0 = planner's choice); used with VK_WGRAD_* variants
        n, kr = int(sys.argv[2]), int(sys.argv[3])
        for c, h in ((96, 128), (192, 64), (288, 32), (64, 128)):
            run(c, h, n, dict(k_rows=kr) if kr else None)
        sys.exit(0)
    n = 16
    for c, h in ((96, 128), (192, 64), (288, 32), (64, 128)):
        run(c, h, n)
        for k_rows in (128, 64):
            for stages in (0, 2):
                run(c, h, n, dict(k_rows=k_rows, stages=stages))
