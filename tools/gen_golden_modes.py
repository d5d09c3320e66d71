"""Fixtures for every constructor-legal configuration of the two wrapper modules beyond the shipped ones
(VERDICT r1 item 7): the UNMODIFIED reference (networks/VIRNet.py) is run on small seeded inputs for each
(extra_mode, noise_cond, kernel_cond, noise_avg) combination; inputs are regenerated from the seed in the tests,
outputs (and, for the trainable-here configurations, the gradient of a fixed linear functional w.r.t. a few
parameters) are stored.

    python tools/gen_golden_modes.py     # needs /root/reference; writes tests/golden/modes.pt
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import ref_import  # noqa: E402

OUT = ROOT / "tests" / "golden" / "modes.pt"
N_FEAT, N_RES, DEP_K = [32, 64, 96], 2, 3

SR_CASES = {
    # name: (ctor kwargs, lr shape, sf)
    "sr_default_down": (dict(), (2, 3, 12, 16), 2),                                   # the class's own defaults
    "sr_input": (dict(extra_mode="Input"), (1, 3, 13, 10), 4),
    "sr_null": (dict(extra_mode="Null"), (1, 3, 12, 12), 2),
    "sr_both_kernel_only": (dict(extra_mode="Both", noise_cond=False), (2, 3, 12, 12), 2),
    "sr_both_noise_only": (dict(extra_mode="Both", kernel_cond=False), (2, 3, 12, 12), 2),
    "sr_null_nocond": (dict(extra_mode="Null", noise_cond=False, kernel_cond=False), (1, 3, 12, 12), 2),
    "sr_both_sigma_map": (dict(extra_mode="Both", noise_avg=False), (2, 3, 11, 14), 2),   # JPEG-noise mode
    "sr_down_sigma_map": (dict(extra_mode="Down", noise_avg=False), (1, 3, 12, 12), 4),
    "sr_input_sigma_map": (dict(extra_mode="Input", noise_avg=False), (1, 3, 12, 12), 2),
    "sr_both_sigma_map_noise_only": (dict(extra_mode="Both", noise_avg=False, kernel_cond=False), (1, 3, 12, 12), 2),
}
DEN_CASES = {
    "den_both": (dict(extra_mode="Both"), (2, 3, 21, 27)),
    "den_down": (dict(extra_mode="Down"), (1, 3, 24, 24)),
    "den_null": (dict(extra_mode="Null"), (1, 3, 24, 24)),
    "den_null_nocond": (dict(extra_mode="Null", noise_cond=False), (1, 3, 24, 24)),
    "den_both_sigma3": (dict(extra_mode="Both", sigma_chn=3), (1, 3, 24, 24)),
}


def sr_kwargs(kw):
    base = dict(im_chn=3, sigma_chn=1, kernel_chn=3, n_feat=N_FEAT, dep_S=5, dep_K=DEP_K, n_resblocks=N_RES)
    base.update(kw)
    return base


def den_kwargs(kw):
    base = dict(im_chn=3, sigma_chn=1, n_feat=N_FEAT, dep_S=5, n_resblocks=N_RES, noise_cond=True, extra_mode="Input",
                noise_avg=False)
    base.update(kw)
    return base


def functional(outs, seed):
    """A fixed linear functional of the outputs (weights regenerated from the seed in the tests)."""
    g = torch.Generator().manual_seed(seed)
    tot = 0.0
    for o in outs:
        tot = tot + (o * torch.randn(o.shape, generator=g) / o.numel() ** 0.5).sum()
    return tot


def main():
    vir, _ = ref_import.import_reference()
    torch.set_num_threads(8)
    out = {}
    for name, (kw, shape, sf) in SR_CASES.items():
        torch.manual_seed(1234)
        net = vir.VIRAttResUNetSR(**sr_kwargs(kw))
        x = torch.rand(*shape, generator=torch.Generator().manual_seed(11))
        mu, kinfo, sigma = net(x, sf)
        functional((mu, kinfo, sigma), 17).backward()
        grads = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        out[name] = dict(mu=mu.detach().clone(), kinfo=kinfo.detach().clone(), sigma=sigma.detach().clone(),
                         grad_norm={k: float(v.norm()) for k, v in grads.items()},
                         grads={k: v for k, v in grads.items() if v.numel() <= 512})
        print(name, tuple(mu.shape), tuple(sigma.shape), float(mu.mean()))
    for name, (kw, shape) in DEN_CASES.items():
        torch.manual_seed(1234)
        net = vir.VIRAttResUNet(**den_kwargs(kw))
        x = torch.rand(*shape, generator=torch.Generator().manual_seed(11))
        mu, sigma = net(x)
        functional((mu, sigma), 17).backward()
        grads = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        out[name] = dict(mu=mu.detach().clone(), sigma=sigma.detach().clone(),
                         grad_norm={k: float(v.norm()) for k, v in grads.items()},
                         grads={k: v for k, v in grads.items() if v.numel() <= 512})
        print(name, tuple(mu.shape), float(mu.mean()))
    torch.save(out, OUT)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
