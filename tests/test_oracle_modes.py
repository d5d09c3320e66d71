"""Every constructor-legal configuration of VIRAttResUNet / VIRAttResUNetSR (extra_mode Null / Input / Down / Both,
noise_cond / kernel_cond off, noise_avg=False sigma maps): the CPU oracle against outputs and gradients of the
UNMODIFIED reference (tests/golden/modes.pt, tools/gen_golden_modes.py).  Runs without a GPU."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
import gen_golden_modes as G  # noqa: E402

from oracle import virnet_oracle as O  # noqa: E402


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def oracle_cfg(kw, sisr):
    k = (G.sr_kwargs if sisr else G.den_kwargs)(kw)
    return O.NetCfg(im_chn=3, sigma_chn=k["sigma_chn"], n_feat=tuple(k["n_feat"]), dep_S=k["dep_S"],
                    dep_K=k.get("dep_K", 8), n_resblocks=k["n_resblocks"], noise_cond=k.get("noise_cond", True),
                    kernel_cond=k.get("kernel_cond", True), extra_mode=k.get("extra_mode", "Down"),
                    noise_avg=k.get("noise_avg", True if sisr else False), sisr=sisr)


@pytest.mark.parametrize("name", list(G.SR_CASES) + list(G.DEN_CASES))
def test_oracle_matches_reference_for_every_configuration(name, golden_dir):
    fx = torch.load(golden_dir / "modes.pt")[name]
    sisr = name in G.SR_CASES
    if sisr:
        kw, shape, sf = G.SR_CASES[name]
    else:
        kw, shape = G.DEN_CASES[name]
    cfg = oracle_cfg(kw, sisr)
    torch.manual_seed(1234)
    sd = {k: v.clone().requires_grad_(True) for k, v in O.build_state_dict(cfg).items()}
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(11))
    outs = O.vir_sisr_forward(sd, x, sf, cfg) if sisr else O.vir_denoise_forward(sd, x, cfg)
    names = ("mu", "kinfo", "sigma") if sisr else ("mu", "sigma")
    for nm, o in zip(names, outs):
        assert o.shape == fx[nm].shape, nm
        assert rel(o.detach(), fx[nm]) < 1e-5, nm
    G.functional(outs, 17).backward()
    for k, gn in fx["grad_norm"].items():
        assert sd[k].grad is not None, k
        assert abs(float(sd[k].grad.norm()) - gn) <= 1e-4 * max(gn, 1e-6) + 1e-7, k
    # parameters the reference leaves without gradient (unused conditioning branches) get none here either
    assert {k for k, v in sd.items() if v.grad is not None} == set(fx["grad_norm"])
