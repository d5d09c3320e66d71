"""GPU bring-up harness for vk_conv_wgrad (one subprocess per case)."""
import argparse
import subprocess
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

CASES = [
    dict(name="wg1x1_bf16_64_64_1tile", dt="bf16", kind="1x1", cin=64, cout=64, n=1, h=8, w=16, tune=dict(ksplit=1)),
    dict(name="wg1x1_bf16_128_64", dt="bf16", kind="1x1", cin=64, cout=128, n=1, h=16, w=16, tune=dict(ksplit=1)),
    dict(name="wg1x1_tf32_64_64", dt="tf32", kind="1x1", cin=64, cout=64, n=1, h=8, w=16, tune=dict(ksplit=1)),
    dict(name="wg3_bf16_c64_1tile", dt="bf16", kind="3x3", cin=64, cout=64, n=1, h=8, w=16, tune=dict(ksplit=1)),
    dict(name="wg3_bf16_c96", dt="bf16", kind="3x3", cin=96, cout=96, n=2, h=32, w=32),
    dict(name="wg3_bf16_c192", dt="bf16", kind="3x3", cin=192, cout=192, n=2, h=32, w=32),
    dict(name="wg3_bf16_c288", dt="bf16", kind="3x3", cin=288, cout=288, n=2, h=16, w=16),
    dict(name="wg3_bf16_c96_ragged", dt="bf16", kind="3x3", cin=96, cout=96, n=1, h=37, w=50),
    dict(name="wg3_tf32_c96", dt="tf32", kind="3x3", cin=96, cout=96, n=2, h=32, w=32),
    dict(name="wg3_tf32_c192_ragged", dt="tf32", kind="3x3", cin=192, cout=192, n=1, h=21, w=27),
    dict(name="wg3_bf16_head_4_96", dt="bf16", kind="3x3", cin=4, cout=96, n=2, h=32, w=32),
    dict(name="wg3_bf16_tail_96_3", dt="bf16", kind="3x3", cin=96, cout=3, n=2, h=32, w=32),
    dict(name="wg3s2_bf16_96_192", dt="bf16", kind="3x3s2", cin=96, cout=192, n=2, h=32, w=32),
    dict(name="wg3s2_tf32_ragged", dt="tf32", kind="3x3s2", cin=96, cout=192, n=1, h=38, w=50),
    dict(name="wgT_bf16_192_96", dt="bf16", kind="convT", cin=192, cout=96, n=2, h=16, w=16),
    dict(name="wgT_tf32_288_192", dt="tf32", kind="convT", cin=288, cout=192, n=1, h=9, w=11),
    dict(name="wg3_bf16_c96_128x128_b4", dt="bf16", kind="3x3", cin=96, cout=96, n=4, h=128, w=128, bench=True),
    dict(name="wg3_bf16_c192_64x64_b4", dt="bf16", kind="3x3", cin=192, cout=192, n=4, h=64, w=64, bench=True),
    dict(name="wg3_bf16_c288_32x32_b4", dt="bf16", kind="3x3", cin=288, cout=288, n=4, h=32, w=32, bench=True),
]


def run_case(idx: int) -> int:
    import torch
    import torch.nn.functional as F
    from virnet_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    c = CASES[idx]
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(100 + idx)
    dtype = ops.VK_BF16 if c["dt"] == "bf16" else ops.VK_TF32

    def q(t):
        if dtype == ops.VK_BF16:
            return t.bfloat16().float()
        return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)

    cin, cout, n, h, w = c["cin"], c["cout"], c["n"], c["h"], c["w"]
    kind = c["kind"]
    x = q(torch.randn(n, cin, h, w, device=dev, generator=g))
    if kind == "convT":
        wt = torch.zeros(cin, cout, 2, 2, device=dev, requires_grad=True)
        bs = torch.zeros(cout, device=dev, requires_grad=True)
        y = F.conv_transpose2d(x, wt, bs, stride=2)
        vk_kind = ops.VK_CONVT2X2_S2
    else:
        k = 1 if kind == "1x1" else 3
        stride = 2 if kind == "3x3s2" else 1
        wt = torch.zeros(cout, cin, k, k, device=dev, requires_grad=True)
        bs = torch.zeros(cout, device=dev, requires_grad=True)
        y = F.conv2d(x, wt, bs, stride=stride, padding=k // 2)
        vk_kind = {"3x3": ops.VK_CONV3X3_S1, "3x3s2": ops.VK_CONV3X3_S2, "1x1": ops.VK_CONV1X1}[kind]
    dy = q(torch.randn(y.shape, device=dev, generator=g))
    y.backward(dy)
    want_w, want_b = wt.grad, bs.grad
    x_nhwc = ops.to_nhwc(x, dtype)
    dy_nhwc = ops.to_nhwc(dy, dtype)
    taps = want_w.shape[2] * want_w.shape[3]
    if kind == "convT":
        a_t, b_t, m_valid, n_valid = x_nhwc, dy_nhwc, cin, cout
    else:
        a_t, b_t, m_valid, n_valid = dy_nhwc, x_nhwc, cout, cin
    ws = torch.zeros(taps, m_valid, n_valid, device=dev)
    db = torch.zeros(m_valid, device=dev)
    tune = c.get("tune")
    ops.conv_wgrad(a_t, b_t, ws, dtype=dtype, kind=vk_kind, m_valid=m_valid, n_valid=n_valid, dbias=db, tune=tune)
    got_w = torch.empty_like(want_w)
    ops.wgrad_unpack(ws, got_w)
    torch.cuda.synchronize()
    ok = True
    tol = 1e-4
    checks = {"dw": (got_w, want_w)}
    if kind != "convT":
        checks["db"] = (db, want_b)
    else:
        checks["db(colsum of x)"] = (db, x.sum(dim=(0, 2, 3)))
    for key, (got, want) in checks.items():
        err = (got - want).abs()
        rel = (err.norm() / want.norm()).item()
        finite = torch.isfinite(got).all().item()
        good = finite and rel < tol
        ok &= good
        print(f"  [{key}] rel_l2={rel:.3e} max_abs={err.max().item():.3e} finite={finite} -> {'OK' if good else 'FAIL'}")
        if not good and key == "dw":
            e = torch.nan_to_num(err, nan=1e3)
            print("   err by tap:", [round(v, 4) for v in e.mean(dim=(0, 1)).flatten().tolist()])
            print("   err by m (first 24, strided):", [round(v, 4) for v in e.mean(dim=(1, 2, 3)).tolist()[:: max(1, e.shape[0] // 24)]])
            print("   err by n (first 24, strided):", [round(v, 4) for v in e.mean(dim=(0, 2, 3)).tolist()[:: max(1, e.shape[1] // 24)]])
            print("   got[0,:6,0,0]:", got[0, :6, 0, 0].tolist())
            print("   want[0,:6,0,0]:", want[0, :6, 0, 0].tolist())
    if c.get("bench") and ok:
        flops = 2.0 * n * h * w * cout * cin * 9
        for tn in (None, dict(ksplit=16), dict(ksplit=32), dict(ksplit=96), dict(k_rows=64), dict(stages=2)):
            try:
                for _ in range(3):
                    ops.conv_wgrad(a_t, b_t, ws, dtype=dtype, kind=vk_kind, m_valid=m_valid, n_valid=n_valid, dbias=db, tune=tn)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                iters = 20
                for _ in range(iters):
                    ops.conv_wgrad(a_t, b_t, ws, dtype=dtype, kind=vk_kind, m_valid=m_valid, n_valid=n_valid, dbias=db, tune=tn)
                e.record()
                torch.cuda.synchronize()
                ms = s.elapsed_time(e) / iters
                print(f"  bench tune={tn}: {ms * 1e3:.1f} us  {flops / ms / 1e9:.1f} TFLOP/s")
            except Exception as ex:  # noqa: BLE001
                print(f"  bench tune={tn}: {ex}")
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
