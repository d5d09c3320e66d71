"""Whole-network parity on the GPU: virnet_b200 (CUDA) vs the CPU oracle, shared weights.

  python tools/debug_net.py [--precision tf32|bf16] [--small]
"""
import argparse
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import torch  # noqa: E402


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--shape", default="2,64,64")
    ap.add_argument("--no-backward", action="store_true")
    args = ap.parse_args()
    import virnet_b200
    from virnet_b200.loss.ELBO_simple import elbo_denoising_simple
    from oracle import virnet_oracle as O

    n_feat = [32, 64, 96] if args.small else [96, 192, 288]
    n_res = 2 if args.small else 3
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=n_feat, dep_S=5, n_resblocks=n_res,
                                    noise_cond=True, extra_mode="Input", noise_avg=False, precision=args.precision)
    cfg = O.NetCfg(n_feat=tuple(n_feat), n_resblocks=n_res)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    n, h, w = (int(v) for v in args.shape.split(","))
    g = torch.Generator().manual_seed(0)
    im_gt = torch.rand(n, 3, h, w, generator=g)
    sig = 5 / 255 + torch.rand(n, 1, h, w, generator=g) * 70 / 255
    im_noisy = im_gt + torch.randn(n, 3, h, w, generator=g) * sig
    sigma_gt = (sig ** 2).clamp_min(1e-10)
    alpha0, eps2 = 24.5, 1e-6

    t0 = time.time()
    (loss_o, lh_o, kg_o, ig_o), mu_o, sg_o, grads_o = O.denoise_loss_and_grads(sd, cfg, im_noisy, im_gt, sigma_gt, alpha0, eps2)
    print(f"oracle: {time.time() - t0:.1f}s loss={loss_o.item():.4f} lh={lh_o.item():.5f} kg={kg_o.item():.3f} ig={ig_o.item():.5f}")

    x = im_noisy.cuda()
    with torch.no_grad():
        mu, sigma = net(x)
    torch.cuda.synchronize()
    print(f"[fwd no_grad] rel(mu)={rel(mu.cpu(), mu_o):.3e} rel(sigma)={rel(sigma.cpu(), sg_o):.3e} "
          f"max|dmu|={(mu.cpu() - mu_o).abs().max().item():.3e} mu.mean={mu.mean().item():.6f}")
    if args.no_backward:
        return
    mu, sigma = net(x)
    beta0 = (alpha0 * sigma_gt).cuda()
    loss, lh, kg, ig = elbo_denoising_simple(mu, sigma, x, im_gt.cuda(), eps2, alpha0, beta0)
    loss.backward()
    torch.cuda.synchronize()
    print(f"[loss] ours: loss={loss.item():.4f} lh={lh.item():.5f} kg={kg.item():.3f} ig={ig.item():.5f}")
    # loss kernel alone, on the oracle's mu/sigma
    from virnet_b200 import ops
    dmu = torch.empty_like(mu)
    dsg = torch.empty_like(sigma)
    out4 = ops.elbo_denoise(mu_o.cuda(), sg_o.cuda(), x, im_gt.cuda(), beta0, eps2=eps2, alpha0=alpha0,
                            digamma_am1=float(torch.digamma(torch.tensor(alpha0 - 1.0, dtype=torch.float64))),
                            d_mu=dmu, d_sigma=dsg)
    muo = mu_o.clone().requires_grad_(True)
    sgo = sg_o.clone().requires_grad_(True)
    lo = O.elbo_denoising_simple(muo, sgo, im_noisy, im_gt, eps2, alpha0, alpha0 * sigma_gt)
    lo[0].backward()
    print(f"[elbo kernel] loss rel={abs(out4[0].item() - lo[0].item()) / abs(lo[0].item()):.2e} "
          f"d_mu rel={rel(dmu.cpu(), muo.grad):.2e} d_sigma rel={rel(dsg.cpu(), sgo.grad):.2e}")
    worst = 0.0
    tot_r = tot_s = 0.0
    tot_ro = tot_so = 0.0
    for name, p in net.named_parameters():
        gr = p.grad.cpu()
        go = grads_o[name]
        r = rel(gr, go)
        worst = max(worst, r)
        if name.startswith("RNet"):
            tot_r += gr.double().pow(2).sum().item(); tot_ro += go.double().pow(2).sum().item()
        else:
            tot_s += gr.double().pow(2).sum().item(); tot_so += go.double().pow(2).sum().item()
        flag = "" if r < (2e-3 if args.precision == "tf32" else 5e-2) else "   <<<<"
        print(f"  grad {name:45s} rel={r:.3e} |g|={go.norm().item():.3e}{flag}")
    print(f"worst rel grad err = {worst:.3e}; grad norms ours R={tot_r ** 0.5:.5e} S={tot_s ** 0.5:.5e} "
          f"oracle R={tot_ro ** 0.5:.5e} S={tot_so ** 0.5:.5e}")


if __name__ == "__main__":
    main()
