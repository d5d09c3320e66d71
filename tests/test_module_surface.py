"""Drop-in surface of the nn.Module classes: constructor signature, attributes,
state_dict keys/shapes and seed-1234 initial values identical to the reference's."""
import pytest
import torch

import virnet_b200
from oracle import virnet_oracle as O
from virnet_b200.lib import VkError


def test_denoise_state_dict_matches_reference_layout(kat):
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNet(im_chn=3, sigma_chn=1, n_feat=[96, 192, 288], dep_S=5, n_resblocks=3,
                                    noise_cond=True, extra_mode="Input", noise_avg=False)
    torch.manual_seed(1234)
    ref = O.build_state_dict(O.NetCfg())
    sd = net.state_dict()
    assert list(sd) == list(ref)
    for k in sd:
        assert sd[k].shape == ref[k].shape and sd[k].dtype == torch.float32
        assert torch.equal(sd[k], ref[k]), k
    assert abs(sum(p.double().sum().item() for p in net.parameters()) - kat["den_syn_128"]["param_sum"]) < 1e-9
    names = [n for n, _ in net.named_parameters()]
    assert all("snet" in n.lower() or "rnet" in n.lower() for n in names)
    assert hasattr(net, "SNet") and hasattr(net, "RNet")
    assert sd["RNet.up_path.0.upsampler.weight"].shape == (288, 192, 2, 2)
    assert len(list(net.buffers())) == 0


def test_sisr_state_dict_matches_reference_layout(kat):
    torch.manual_seed(1234)
    net = virnet_b200.VIRAttResUNetSR(im_chn=3, sigma_chn=1, dep_S=5, dep_K=8, n_feat=[96, 160, 224], n_resblocks=2,
                                      extra_mode="Both", noise_avg=True, noise_cond=True, kernel_cond=True)
    torch.manual_seed(1234)
    ref = O.build_state_dict(O.NetCfg(n_feat=(96, 160, 224), n_resblocks=2, extra_mode="Both", noise_avg=True, sisr=True))
    sd = net.state_dict()
    assert list(sd) == list(ref) and len(sd) == 225
    assert all(torch.equal(sd[k], ref[k]) for k in sd)
    assert "KNet.head.bias" not in sd and sd["KNet.head.weight"].shape == (64, 3, 9, 9)


def test_defaults_and_asserts_follow_the_reference():
    net = virnet_b200.VIRAttResUNet(3)
    assert net.SNet.conv_last.out_channels == 3 and net.RNet.n_feat == [64, 128, 192]
    with pytest.raises(AssertionError):
        virnet_b200.VIRAttResUNet(3, extra_mode="sideways")
    with pytest.raises(AssertionError):
        virnet_b200.VIRAttResUNet(3, n_feat=64)
    virnet_b200.VIRAttResUNet(3, extra_mode="NULL")          # case-insensitive


def test_no_cpu_fallback():
    net = virnet_b200.VIRAttResUNet(3, sigma_chn=1, n_feat=[32, 64], n_resblocks=1)
    with pytest.raises(VkError):
        net(torch.rand(1, 3, 16, 16))


def test_load_state_dict_roundtrip_with_module_prefix():
    torch.manual_seed(0)
    a = virnet_b200.VIRAttResUNet(3, sigma_chn=1, n_feat=[32, 64], n_resblocks=1)
    b = virnet_b200.VIRAttResUNet(3, sigma_chn=1, n_feat=[32, 64], n_resblocks=1)
    ddp_style = {"module." + k: v for k, v in a.state_dict().items()}
    b.load_state_dict({k[7:]: v for k, v in ddp_style.items()}, strict=True)   # scripts/testing_demo.py:69-72
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
