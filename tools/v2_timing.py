#!/usr/bin/env python
"""Per-role stall breakdown of the persistent conv kernel (vk_conv_v2.cuh) at the bench layer shapes.
Uses vk_conv_args.cta_timing (int64[grid][16]); debug / tuning aid."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from virnet_b200 import ops  # noqa: E402
from tools.conv_bench import make, time_it  # noqa: E402

NAMES = ["total", "mma:wait_A", "mma:wait_B", "mma:wait_tmem", "mma:span", "A:wait_empty", "B0:wait_empty",
         "epi0:wait_tmem_full", "epi0:wait_in", "epi0:wait_store_read", "epi0:span", "A:span", "B0:span",
         "epi0:tmem_ld", "epi0:math+sts", "epi0:group_sync", "epi0:store_issue", "epi0:misc(in-wait,lds,coords)", "epi0:fence"]


def run(c_in, c_out, h, n, epi, dt=ops.VK_BF16, kind=ops.VK_CONV3X3_S1, tune=None, impl=4):
    x, w, kw, o1, o2 = make(c_in, c_out, h, n, epi, dt, kind)
    t = dict(tune or {})
    t["impl"] = impl
    us = time_it(lambda: ops.conv_igemm(x, w, tune=t, **kw))
    timing = torch.zeros(148 * 24, device="cuda", dtype=torch.int64)
    t["cta_timing"] = timing
    ops.conv_igemm(x, w, tune=t, **kw)
    torch.cuda.synchronize()
    tm = timing.view(-1, 24).cpu().float()
    tm = tm[tm[:, 0] != 0]
    oh = o1.shape[1]
    flops = 2.0 * n * oh * oh * 9 * c_in * c_out
    m = tm.mean(0)
    print(f"impl={impl} Cin={c_in} Cout={c_out} {h}x{h} n={n} epi={epi} tune={tune}: {us:.1f} us {flops / us / 1e6:.0f} TF/s | "
          + " ".join(f"{NAMES[i]}={m[i]:.0f}" for i in range(len(NAMES))) + f" | max total={tm[:, 0].max():.0f}", flush=True)


if __name__ == "__main__":
    n = 16
    for epi in ("out2", "resid_dual"):
        run(96, 96, 128, n, epi)
    run(192, 192, 64, n, "resid_dual")
