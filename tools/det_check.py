#!/usr/bin/env python
"""Bitwise repeatability of one forward + loss + backward (deterministic mode): runs the same step several times on the
same weights and inputs and compares every activation / gradient buffer of the engine with the first run."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import bench  # noqa: E402
import virnet_b200  # noqa: E402
from virnet_b200 import ops  # noqa: E402


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    dev = torch.device("cuda")
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=bench.N_FEAT, dep_S=bench.DEP_S, n_resblocks=bench.N_RES,
                                    noise_cond=True, extra_mode="Input", noise_avg=False, precision=prec).to(dev)
    eng = net.engine()
    eng.deterministic = True
    eng._ensure_flat()
    x, gt, sg = bench.synth_batch(b, 0, dev)
    d_mu = d_sigma = None
    ref = None
    acc, out4 = ops.elbo_ws(dev), torch.zeros(4, device=dev)
    for it in range(iters):
        mu, sigma = eng.forward(x, save=True)
        if d_mu is None:
            d_mu, d_sigma = torch.empty_like(mu), torch.empty_like(sigma)
        ops.elbo_denoise(mu, sigma, x, gt, sg, beta0_scale=bench.ALPHA0, eps2=bench.EPS2, alpha0=bench.ALPHA0,
                         digamma_am1=0.0, d_mu=d_mu, d_sigma=d_sigma, acc3=acc, out4=out4)
        st = eng.saved["set"]
        eng.backward(d_mu, d_sigma)
        torch.cuda.synchronize()
        snap = {"mu": mu.clone(), "sigma": sigma.clone(), "d_mu": d_mu.clone(), "loss": out4.clone(),
                "flat_grads": eng.flat_grads.clone()}
        for (name, shape, dt), t in st["bufs"].items():
            snap[f"buf:{name}"] = t.clone()
        for i, ly in enumerate(eng.layers):
            if ly.det is not None:
                snap[f"partials:{i}:{ly.name}"] = ly.det["partials"].clone()
        if ref is None:
            ref = snap
            print(f"run 0: {len(snap)} tensors recorded, loss {out4[0].item():.6f}", flush=True)
            continue
        bad = []
        for k, v in snap.items():
            r = ref[k]
            if not torch.equal(v.view(torch.uint8), r.view(torch.uint8)):
                neq = (v.view(torch.uint8) != r.view(torch.uint8))
                idx = neq.flatten().nonzero().flatten()
                bad.append((k, tuple(v.shape), int(idx.numel()), int(idx[0]), int(idx[-1])))
        print(f"run {it}: {len(bad)} tensors differ", flush=True)
        for rec in bad[:40]:
            print("   ", rec, flush=True)


if __name__ == "__main__":
    main()
