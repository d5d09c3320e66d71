#!/usr/bin/env python
"""Times vk_conv_igemm at the bench layer shapes for both kernel generations (force_impl 1 = v1,
2 = v2) and cross-checks their outputs.  Debug / tuning aid; prints one line per case."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from virnet_b200 import ops  # noqa: E402

dev = "cuda"


def make(c_in, c_out, h, n, epi, dt, kind=ops.VK_CONV3X3_S1):
    tdt = ops.TORCH_DTYPE[dt]
    x = torch.randn(n, h, h, c_in, device=dev).to(tdt)
    taps = 9
    w = (torch.randn(taps, c_out, c_in, device=dev) / (3 * c_in ** 0.5)).to(tdt)
    bias = torch.randn(c_out, device=dev)
    oh = h if kind == ops.VK_CONV3X3_S1 else (h + 1) // 2
    o1 = torch.zeros(n, oh, oh, c_out, device=dev, dtype=tdt)
    o2 = torch.zeros_like(o1)
    resid = torch.randn_like(o1)
    mask = torch.randn_like(o1)
    kw = dict(dtype=dt, kind=kind, cout=c_out, bias=bias, ldo=c_out, alpha=0.2)
    if epi == "out2":
        kw.update(out2=o2)
    elif epi == "dual":
        kw.update(out1=o1, out2=o2)
    elif epi == "resid_dual":
        kw.update(resid=resid, out1=o1, out2=o2)
    elif epi == "mask":
        kw.update(mask=mask, out1=o1)
    elif epi == "mask_resid":
        kw.update(resid=resid, mask=mask, out1=o1)
    return x, w, kw, o1, o2


def time_it(fn, iters=20):
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3  # us


def run(c_in, c_out, h, n, epi, dt=ops.VK_BF16, kind=ops.VK_CONV3X3_S1, tune=None, check=True, impls=(1, 3, 4)):
    x, w, kw, o1, o2 = make(c_in, c_out, h, n, epi, dt, kind)
    res = {}
    oh = o1.shape[1]
    flops = 2.0 * n * oh * oh * 9 * c_in * c_out
    line = f"Cin={c_in} Cout={c_out} {h}x{h} n={n} epi={epi} {'bf16' if dt == ops.VK_BF16 else 'tf32'} tune={tune}:"
    for impl in impls:
        t = dict(tune or {})
        t["impl"] = impl
        if impl == 1:
            t.pop("p", None), t.pop("stages", None)
        o1.zero_(), o2.zero_()
        try:
            ops.conv_igemm(x, w, tune=t, **kw)
            torch.cuda.synchronize()
        except Exception as ex:  # noqa: BLE001
            line += f" v{impl}: {ex};"
            continue
        res[impl] = (o1.float().clone(), o2.float().clone())
        us = time_it(lambda: ops.conv_igemm(x, w, tune=t, **kw))
        line += f" v{impl}: {us:7.1f} us {flops / us / 1e6:6.0f} TF/s;"
    if check and len(res) > 1:
        first = impls[0]
        for impl in impls[1:]:
            if impl not in res or first not in res:
                continue
            errs = []
            for a, b in zip(res[first], res[impl]):
                den = a.abs().max().item() or 1.0
                errs.append((a - b).abs().max().item() / den)
            line += f" diff v{first}/v{impl} = {max(errs):.2e}"
    print(line, flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    n = args.n
    B = ops.VK_BF16
    run(96, 96, 128, n, "out2")
    run(96, 96, 128, n, "resid_dual")
    if not args.quick:
        run(96, 96, 128, n, "mask")
        run(96, 96, 128, n, "mask_resid")
        for p in (1, 2):
            run(96, 96, 128, n, "out2", tune=dict(p=p), check=False)
    run(192, 192, 64, n, "out2")
    run(192, 192, 64, n, "resid_dual")
    run(288, 288, 32, n, "out2")
    run(288, 288, 32, n, "resid_dual")
    run(64, 64, 128, n, "out2")
    if not args.quick:
        run(96, 192, 128, n, "dual", kind=ops.VK_CONV3X3_S2)
        run(192, 288, 64, n, "dual", kind=ops.VK_CONV3X3_S2)
        run(16, 96, 128, n, "dual")
        run(96, 96, 128, n, "resid_dual", dt=ops.VK_TF32)
        run(192, 192, 64, n, "resid_dual", dt=ops.VK_TF32)
