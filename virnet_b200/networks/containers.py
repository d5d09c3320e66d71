"""Parameter containers with the reference's module tree.

The reference's trainers and scripts rely on the module *surface* only:
`state_dict()` keys and OIHW fp32 shapes, `named_parameters()` names containing
snet / rnet / knet, `.SNet/.RNet/.KNet` attributes, and — under `torch.manual_seed` — the
parameter creation order, which fixes the initial values.  These classes reproduce that
surface (reference: networks/DnCNN.py:8-52, networks/KNet.py:12-59,
networks/AttResUNet.py:11-139) but hold no arithmetic: every op of the forward and
backward pass runs in libvirnet_sm100.so, driven by virnet_b200/engine.py.
Calling a container directly is therefore an error.
"""
from __future__ import annotations

import torch
from torch import nn


def _no_direct_call(self, *a, **k):
    raise RuntimeError(
        f"{type(self).__name__} is a parameter container of virnet_b200; run the enclosing "
        "VIRAttResUNet / VIRAttResUNetSR module instead (the whole network executes as one CUDA program)")


class _Container(nn.Module):
    forward = _no_direct_call


def _conv3(cin, cout, stride=1, bias=True):
    return nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1, bias=bias)


def _conv1(cin, cout):
    return nn.Conv2d(cin, cout, kernel_size=1, stride=1, padding=0)


class DnCNN(_Container):
    """SNet: `dep` 3x3 convs, 64 filters, LeakyReLU(0.25); orthogonal init, zero bias."""

    def __init__(self, in_channels, out_channels, dep=5, num_filters=64, noise_avg=False):
        super().__init__()
        self.dep = dep
        self.noise_avg = noise_avg
        self.conv1 = _conv3(in_channels, num_filters)
        self.relu = nn.LeakyReLU(0.25, True)
        mid = []
        for _ in range(dep - 2):
            mid += [_conv3(num_filters, num_filters), nn.LeakyReLU(0.25, True)]
        self.mid_layer = nn.Sequential(*mid)
        self.conv_last = _conv3(num_filters, out_channels)
        self.global_avg = nn.AdaptiveAvgPool2d((1, 1)) if noise_avg else nn.Identity()
        gain = nn.init.calculate_gain("leaky_relu", 0.25)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.orthogonal_(m.weight, gain=gain)
                nn.init.zeros_(m.bias)

    def conv_layers(self):
        return [self.conv1] + [m for m in self.mid_layer if isinstance(m, nn.Conv2d)] + [self.conv_last]


class CALayer(_Container):
    def __init__(self, nf, reduction=16):
        super().__init__()
        self.avg = nn.AdaptiveAvgPool2d(1)
        self.body = nn.Sequential(_conv1(nf, nf // reduction), nn.LeakyReLU(0.2),
                                  _conv1(nf // reduction, nf), nn.Sigmoid())


class RB_Layer(_Container):
    def __init__(self, nf):
        super().__init__()
        self.body = nn.Sequential(_conv3(nf, nf), nn.LeakyReLU(0.2, True), _conv3(nf, nf), CALayer(nf))


class KernelNet(_Container):
    """KNet: 9x9 stride-4 head (no bias), `num_blocks` channel-attention residual blocks, 3x3 tail + GAP."""

    def __init__(self, in_nc=3, out_chn=3, nf=64, num_blocks=8, scale=4):
        super().__init__()
        self.head = nn.Conv2d(in_nc, nf, kernel_size=9, stride=4, padding=4, bias=False)
        self.body = nn.Sequential(*[RB_Layer(nf) for _ in range(num_blocks)])
        self.tail = nn.Sequential(_conv3(nf, out_chn), nn.AdaptiveAvgPool2d((1, 1)))


class AttLayer(_Container):
    """SFT-style modulation MLP of 1x1 convs: extra -> C/8 -> C/4 -> {mul (sigmoid), add}."""

    def __init__(self, out_chn=64, extra_chn=4):
        super().__init__()
        nf1, nf2 = out_chn // 8, out_chn // 4
        self.conv1 = _conv1(extra_chn, nf1)
        self.leaky1 = nn.LeakyReLU(0.2)
        self.conv2 = _conv1(nf1, nf2)
        self.leaky2 = nn.LeakyReLU(0.2)
        self.mul_conv = _conv1(nf2, out_chn)
        self.sig = nn.Sigmoid()
        self.add_conv = _conv1(nf2, out_chn)


class AttResBlock(_Container):
    def __init__(self, nf=64, extra_chn=4):
        super().__init__()
        self.extra_chn = extra_chn
        if extra_chn > 0:
            self.sft1 = AttLayer(nf, extra_chn)
            self.sft2 = AttLayer(nf, extra_chn)
        self.lrelu1 = nn.LeakyReLU(0.2)
        self.conv1 = _conv3(nf, nf)
        self.lrelu2 = nn.LeakyReLU(0.2)
        self.conv2 = _conv3(nf, nf)


class DownBlock(_Container):
    def __init__(self, in_chn=64, out_chn=128, extra_chn=4, n_resblocks=1, downsample=True):
        super().__init__()
        self.body = nn.ModuleList([AttResBlock(in_chn, extra_chn) for _ in range(n_resblocks)])
        self.downsampler = _conv3(in_chn, out_chn, stride=2) if downsample else nn.Identity()


class UpBlock(_Container):
    def __init__(self, in_chn=128, out_chn=64, n_resblocks=1):
        super().__init__()
        self.upsampler = nn.ConvTranspose2d(in_chn, out_chn, kernel_size=2, stride=2, padding=0)
        self.body = nn.ModuleList([AttResBlock(nf=out_chn, extra_chn=0) for _ in range(n_resblocks)])


class AttResUNet(_Container):
    """RNet: U-Net of pre-activation residual blocks (strided-conv down, ConvT up)."""

    def __init__(self, in_chn=3, extra_chn=4, out_chn=3, n_resblocks=2, n_feat=(64, 128, 196, 256),
                 extra_mode="Input"):
        super().__init__()
        assert isinstance(n_feat, (tuple, list))
        self.depth = len(n_feat)
        self.n_feat = list(n_feat)
        self.n_resblocks = n_resblocks
        self.in_chn, self.extra_chn, self.out_chn = in_chn, extra_chn, out_chn
        self.extra_mode = extra_mode.lower()
        assert self.extra_mode in ("null", "input", "down", "both")
        head_in = in_chn if self.extra_mode in ("down", "null") else in_chn + extra_chn
        self.head = _conv3(head_in, n_feat[0])
        extra_down = extra_chn if self.extra_mode in ("down", "both") else 0
        self.down_path = nn.ModuleList()
        for ii in range(self.depth):
            last = ii + 1 == self.depth
            self.down_path.append(DownBlock(n_feat[ii], n_feat[ii] if last else n_feat[ii + 1], extra_chn=extra_down,
                                            n_resblocks=n_resblocks, downsample=not last))
        self.up_path = nn.ModuleList()
        for jj in reversed(range(self.depth - 1)):
            self.up_path.append(UpBlock(n_feat[jj + 1], n_feat[jj], n_resblocks))
        self.tail = _conv3(n_feat[0], out_chn)
