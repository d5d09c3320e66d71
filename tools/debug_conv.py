"""GPU bring-up harness for vk_conv_igemm: each case runs in its own process so
that a trapped kernel cannot poison the next one.

  python tools/debug_conv.py            # run every case (subprocess per case)
  python tools/debug_conv.py --case 3   # one case, in-process
"""
import argparse
import subprocess
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

CASES = [
    # name, dtype, kind, cin, cout, n, h, w, extras
    dict(name="gemm1x1_bf16_c64_chunk128", dt="bf16", kind="1x1", cin=64, cout=64, n=1, h=8, w=16, tune=dict(chunk=128)),
    dict(name="gemm1x1_bf16_c64_chunk64", dt="bf16", kind="1x1", cin=64, cout=64, n=1, h=8, w=16, tune=dict(chunk=64)),
    dict(name="gemm1x1_bf16_c64_chunk32", dt="bf16", kind="1x1", cin=64, cout=64, n=1, h=8, w=16, tune=dict(chunk=32)),
    dict(name="gemm1x1_tf32_c64_chunk128", dt="tf32", kind="1x1", cin=64, cout=64, n=1, h=8, w=16, tune=dict(chunk=128)),
    dict(name="conv3_bf16_c64_1tile", dt="bf16", kind="3x3", cin=64, cout=64, n=1, h=8, w=16, tune=dict(chunk=128, p=1)),
    dict(name="conv3_bf16_c96_32x32", dt="bf16", kind="3x3", cin=96, cout=96, n=2, h=32, w=32),
    dict(name="conv3_bf16_c96_p1", dt="bf16", kind="3x3", cin=96, cout=96, n=2, h=32, w=32, tune=dict(p=1)),
    dict(name="conv3_bf16_c192", dt="bf16", kind="3x3", cin=192, cout=192, n=2, h=32, w=32),
    dict(name="conv3_bf16_c288", dt="bf16", kind="3x3", cin=288, cout=288, n=2, h=32, w=32),
    dict(name="conv3_bf16_c96_ragged", dt="bf16", kind="3x3", cin=96, cout=96, n=1, h=37, w=50),
    dict(name="conv3_tf32_c96", dt="tf32", kind="3x3", cin=96, cout=96, n=2, h=32, w=32),
    dict(name="conv3_tf32_c192_ragged", dt="tf32", kind="3x3", cin=192, cout=192, n=1, h=21, w=27),
    dict(name="conv3_bf16_head_c16_96", dt="bf16", kind="3x3", cin=4, cout=96, n=2, h=32, w=32),
    dict(name="conv3_bf16_tail_96_3_nchw", dt="bf16", kind="3x3", cin=96, cout=3, n=2, h=32, w=32, epi="nchw"),
    dict(name="conv3_bf16_epi_full", dt="bf16", kind="3x3", cin=96, cout=96, n=2, h=32, w=32, epi="full"),
    dict(name="conv3_tf32_epi_full", dt="tf32", kind="3x3", cin=96, cout=96, n=2, h=32, w=32, epi="full"),
    dict(name="conv3s2_bf16_96_192", dt="bf16", kind="3x3s2", cin=96, cout=192, n=2, h=32, w=32),
    dict(name="conv3s2_bf16_ragged", dt="bf16", kind="3x3s2", cin=96, cout=192, n=1, h=38, w=50),
    dict(name="convT_bf16_192_96", dt="bf16", kind="convT", cin=192, cout=96, n=2, h=16, w=16, epi="resid"),
    dict(name="convT_tf32_288_192", dt="tf32", kind="convT", cin=288, cout=192, n=1, h=9, w=11),
    dict(name="conv2x2s2_bf16_96_192", dt="bf16", kind="2x2s2", cin=96, cout=192, n=2, h=32, w=32),
    dict(name="conv2x2s2_tf32_192_288", dt="tf32", kind="2x2s2", cin=192, cout=288, n=1, h=20, w=24),
    dict(name="s2dgrad_bf16_192_96", dt="bf16", kind="3x3s2dgrad", cin=192, cout=96, n=2, h=32, w=32, epi="resid"),
    dict(name="s2dgrad_tf32_odd", dt="tf32", kind="3x3s2dgrad", cin=192, cout=96, n=1, h=37, w=51),
    dict(name="conv3_bf16_c96_128x128_b4", dt="bf16", kind="3x3", cin=96, cout=96, n=4, h=128, w=128, bench=True),
    dict(name="conv3_bf16_c192_64x64_b4", dt="bf16", kind="3x3", cin=192, cout=192, n=4, h=64, w=64, bench=True),
    dict(name="conv3_bf16_c288_32x32_b4", dt="bf16", kind="3x3", cin=288, cout=288, n=4, h=32, w=32, bench=True),
    dict(name="conv3_tf32_c96_128x128_b4", dt="tf32", kind="3x3", cin=96, cout=96, n=4, h=128, w=128, bench=True),
]


def run_case(idx: int) -> int:
    import torch
    import torch.nn.functional as F
    from virnet_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    c = CASES[idx]
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(idx)
    dtype = ops.VK_BF16 if c["dt"] == "bf16" else ops.VK_TF32

    def q(t):  # quantise to the operand format so the fp32 reference sees identical operands
        if dtype == ops.VK_BF16:
            return t.bfloat16().float()
        return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)

    cin, cout, n, h, w = c["cin"], c["cout"], c["n"], c["h"], c["w"]
    kind = c["kind"]
    x = q(torch.randn(n, cin, h, w, device=dev, generator=g))
    bias = torch.randn(cout, device=dev, generator=g) * 0.5
    epi = c.get("epi", "plain")
    tune = c.get("tune")
    x_nhwc = ops.to_nhwc(x, dtype)
    ldx = x_nhwc.shape[-1]

    if kind in ("3x3", "3x3s2", "1x1"):
        k = 1 if kind == "1x1" else 3
        wt = q(torch.randn(cout, cin, k, k, device=dev, generator=g) / (cin * k * k) ** 0.5)
        stride = 2 if kind == "3x3s2" else 1
        ref = F.conv2d(x, wt, bias, stride=stride, padding=k // 2)
        wp = ops.pack_conv_weight(wt, dtype, ldx)
        vk_kind = {"3x3": ops.VK_CONV3X3_S1, "3x3s2": ops.VK_CONV3X3_S2, "1x1": ops.VK_CONV1X1}[kind]
        wrows = wp.shape[1]
    elif kind == "2x2s2":
        wt = q(torch.randn(cout, cin, 2, 2, device=dev, generator=g) / (cin * 4) ** 0.5)
        ref = F.conv2d(x, wt, bias, stride=2)
        wp = ops.pack_conv_weight(wt, dtype, ldx)
        vk_kind = ops.VK_CONV2X2_S2
        wrows = wp.shape[1]
    elif kind == "3x3s2dgrad":
        # x plays dY of a stride-2 conv (cout_fwd = cin here) whose input had `cout` channels and size (h, w)
        wt = q(torch.randn(cin, cout, 3, 3, device=dev, generator=g) / (cin * 9) ** 0.5)   # [Co_f, Ci_f, 3, 3]
        fine = torch.zeros(n, cout, h, w, device=dev, requires_grad=True)
        yf = F.conv2d(fine, wt, None, stride=2, padding=1)
        x = q(torch.randn(yf.shape, device=dev, generator=g))
        yf.backward(x)
        ref = fine.grad + bias.view(1, -1, 1, 1)
        x_nhwc = ops.to_nhwc(x, dtype)
        ldx = x_nhwc.shape[-1]
        # packed [t][ci_f][co_f] = W[co_f][ci_f][t]  (transposed, not rotated)
        wp = ops.pack_conv_weight(wt.permute(1, 0, 2, 3).contiguous(), dtype, ldx)
        vk_kind = ops.VK_CONV3X3_S2_DGRAD
        wrows = wp.shape[1]
    else:
        wt = q(torch.randn(cin, cout, 2, 2, device=dev, generator=g) / (cin) ** 0.5)
        ref = F.conv_transpose2d(x, wt, bias, stride=2)
        wp = ops.pack_convT_weight(wt, dtype, ldx)
        vk_kind = ops.VK_CONVT2X2_S2
        wrows = wp.shape[1]
    bias_p = bias.clone()
    oh, ow = ref.shape[-2:]
    ldo = ops.chan_pad(cout, dtype)
    tdt = ops.TORCH_DTYPE[dtype]
    results = {}

    if epi == "nchw":
        xin = torch.randn(n, cout, oh - 1, ow - 2, device=dev, generator=g)
        out = torch.full((n, cout, oh - 1, ow - 2), float("nan"), device=dev)
        ops.conv_igemm(x_nhwc, wp, dtype=dtype, kind=vk_kind, cout=cout, bias=bias_p, epi=ops.VK_EPI_NCHW_F32,
                       resid=xin, out1=out, crop=(oh - 1, ow - 2), tune=tune)
        results["nchw"] = (out, ref[..., : oh - 1, : ow - 2] + xin)
        out2 = torch.full((n, cout, oh, ow), float("nan"), device=dev)
        ops.conv_igemm(x_nhwc, wp, dtype=dtype, kind=vk_kind, cout=cout, bias=bias_p, epi=ops.VK_EPI_NCHW_F32,
                       out1=out2, act_expclamp=True, clamp=(-1.0, 0.5), tune=tune)
        results["expclamp"] = (out2, torch.exp(torch.clamp(ref, -1.0, 0.5)))
    elif epi in ("full", "resid"):
        resid = q(torch.randn(n, cout, oh, ow, device=dev, generator=g))
        maskt = q(torch.randn(n, cout, oh, ow, device=dev, generator=g))
        o1 = torch.full((n, oh, ow, ldo), float("nan"), device=dev, dtype=tdt)
        o2 = torch.full((n, oh, ow, ldo), float("nan"), device=dev, dtype=tdt)
        use_mask = epi == "full"
        ops.conv_igemm(x_nhwc, wp, dtype=dtype, kind=vk_kind, cout=cout, bias=bias_p, ldo=ldo,
                       resid=ops.to_nhwc(resid, dtype, ldo), mask=ops.to_nhwc(maskt, dtype, ldo) if use_mask else None,
                       out1=o1, out2=o2, alpha=0.2, tune=tune, out_hw=(oh, ow))
        v = ref
        if use_mask:
            v = v * torch.where(maskt > 0, 1.0, 0.2)
        v = v + resid
        results["out1"] = (ops.from_nhwc(o1, cout), v)
        results["out2"] = (ops.from_nhwc(o2, cout), F.leaky_relu(v, 0.2))
    else:
        o1 = torch.full((n, oh, ow, ldo), float("nan"), device=dev, dtype=tdt)
        ops.conv_igemm(x_nhwc, wp, dtype=dtype, kind=vk_kind, cout=cout, bias=bias_p, ldo=ldo, out1=o1, tune=tune, out_hw=(oh, ow))
        results["out1"] = (ops.from_nhwc(o1, cout), ref)
    torch.cuda.synchronize()

    ok = True
    tol = 5e-3 if dtype == ops.VK_BF16 else 2e-4   # bf16 tolerance covers the bf16 OUTPUT rounding
    for key, (got, want) in results.items():
        err = (got - want).abs()
        finite = torch.isfinite(got).all().item()
        rel = (err.norm() / want.norm()).item()
        mx = err.max().item()
        good = finite and rel < tol
        ok &= good
        print(f"  [{key}] rel_l2={rel:.3e} max_abs={mx:.3e} finite={finite} -> {'OK' if good else 'FAIL'}")
        if not good:
            e = torch.nan_to_num(err, nan=1e3)
            print("   err by channel:", [round(v, 3) for v in e.mean(dim=(0, 2, 3)).tolist()[:: max(1, cout // 24)]])
            print("   err by row:", [round(v, 3) for v in e.mean(dim=(0, 1, 3)).tolist()[:24]])
            print("   err by col:", [round(v, 3) for v in e.mean(dim=(0, 1, 2)).tolist()[:24]])
            print("   got[0,:4,0,:6]:", got[0, :4, 0, :6].tolist())
            print("   want[0,:4,0,:6]:", want[0, :4, 0, :6].tolist())

    if c.get("bench") and ok:
        o1 = torch.empty((n, oh, ow, ldo), device=dev, dtype=tdt)
        flops = 2.0 * n * oh * ow * cout * cin * 9
        for tn in (None, dict(p=1), dict(p=2), dict(chunk=32), dict(chunk=128), dict(tw=8), dict(tw=32)):
            t2 = dict(tune or {})
            t2.update(tn or {})
            try:
                for _ in range(3):
                    ops.conv_igemm(x_nhwc, wp, dtype=dtype, kind=vk_kind, cout=cout, bias=bias_p, ldo=ldo, out2=o1, tune=t2)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                iters = 20
                for _ in range(iters):
                    ops.conv_igemm(x_nhwc, wp, dtype=dtype, kind=vk_kind, cout=cout, bias=bias_p, ldo=ldo, out2=o1, tune=t2)
                e.record()
                torch.cuda.synchronize()
                ms = s.elapsed_time(e) / iters
                print(f"  bench tune={t2}: {ms * 1e3:.1f} us  {flops / ms / 1e9:.1f} TFLOP/s")
            except Exception as ex:  # noqa: BLE001
                print(f"  bench tune={t2}: {ex}")
    return 0 if ok else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=-1)
    ap.add_argument("--only", type=str, default="")
    args = ap.parse_args()
    if args.case >= 0:
        print(f"case {args.case}: {CASES[args.case]['name']}", flush=True)
        sys.exit(run_case(args.case))
    fails = []
    for i, c in enumerate(CASES):
        if args.only and args.only not in c["name"]:
            continue
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--case", str(i)], capture_output=True, text=True, timeout=300)
            out, code = r.stdout + r.stderr[-3000:], r.returncode
        except subprocess.TimeoutExpired as ex:
            out, code = f"TIMEOUT {ex}", -9
        print(out.rstrip())
        print(f"=> {c['name']}: {'PASS' if code == 0 else 'FAIL(%d)' % code} ({time.time() - t0:.1f}s)", flush=True)
        if code != 0:
            fails.append(c["name"])
    print("FAILED:", fails)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
