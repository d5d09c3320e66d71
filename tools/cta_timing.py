#!/usr/bin/env python
"""Per-CTA cycle breakdown of vk_conv_igemm at the bench shapes (debug aid; uses vk_conv_args.cta_timing)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from virnet_b200 import ops  # noqa: E402

dev = "cuda"
dt = ops.VK_BF16


def run(c, h, n, epi, tune=None, impl=0):
    x = torch.randn(n, h, h, c, device=dev).bfloat16()
    w = (torch.randn(9, c, c, device=dev) / (3 * c ** 0.5)).bfloat16()
    bias = torch.randn(c, device=dev)
    o1 = torch.empty(n, h, h, c, device=dev, dtype=torch.bfloat16)
    o2 = torch.empty_like(o1)
    resid = torch.randn_like(o1)
    timing = torch.zeros(4096 * 8, device=dev, dtype=torch.int64)
    kw = dict(dtype=dt, kind=ops.VK_CONV3X3_S1, cout=c, bias=bias, ldo=c, alpha=0.2)
    if epi == "out2":
        kw.update(out2=o2)
    elif epi == "resid_dual":
        kw.update(resid=resid, out1=o1, out2=o2)
    elif epi == "mask_resid":
        kw.update(resid=resid, mask=o2, out1=o1)
    t = dict(tune or {})
    t["impl"] = impl or 1   # this tool reads the v1 per-CTA counters (v2: tools/v2_timing.py)
    for _ in range(3):
        ops.conv_igemm(x, w, tune=t, **kw)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.conv_igemm(x, w, tune=t, **kw)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 100
    t["cta_timing"] = timing
    ops.conv_igemm(x, w, tune=t, **kw)
    torch.cuda.synchronize()
    tm = timing.view(-1, 8).cpu()
    tm = tm[tm[:, 0] != 0]
    tot = (tm[:, 2] - tm[:, 0]).float()
    main = (tm[:, 1] - tm[:, 0]).float()
    epi_t = (tm[:, 2] - tm[:, 1]).float()
    issue_end = (tm[:, 5] - tm[:, 0]).float()
    flops = 2.0 * n * h * h * 9 * c * c
    print(f"C={c} {h}x{h} n={n} epi={epi} tune={tune}: {us:.1f} us {flops / us / 1e6:.0f} TFLOP/s | CTAs={len(tm)} "
          f"cyc/CTA total={tot.mean():.0f} mainloop={main.mean():.0f} epilogue={epi_t.mean():.0f} "
          f"prod_wait={tm[:, 3].float().mean():.0f} mma_wait={tm[:, 4].float().mean():.0f} mma_issue_end={issue_end.mean():.0f} "
          f"span={(tm[:, 2].max() - tm[:, 0].min()).item()} CTAs/SM max={torch.bincount(tm[:, 6]).max().item()}")


if __name__ == "__main__":
    for epi in ("out2", "resid_dual"):
        run(96, 128, 16, epi)
    run(96, 128, 16, "out2", dict(p=1))
    run(96, 128, 16, "out2", dict(p=2))
    run(192, 64, 16, "out2")
    run(288, 32, 16, "out2")
    run(64, 128, 16, "out2")
