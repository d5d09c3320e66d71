"""Drop-in for utils/util_net.py:27-65 `forward_chop`: quadrant tiling of very large images with `shave` pixels of
overlap.  The four overlapping quadrants have the same size, so they run as ONE batched forward (the reference runs
them one by one, n_GPUs = 1); its recursive branch (which the reference calls without the `net` argument and therefore
cannot execute) recurses properly here."""
from __future__ import annotations

import torch


def forward_chop(net, x, scale=1, shave=10, min_size=160000):
    b, c, h, w = x.size()
    h_half, w_half = h // 2, w // 2
    h_size, w_size = h_half + shave, w_half + shave
    lr_list = [x[:, :, 0:h_size, 0:w_size], x[:, :, 0:h_size, (w - w_size):w],
               x[:, :, (h - h_size):h, 0:w_size], x[:, :, (h - h_size):h, (w - w_size):w]]
    if w_size * h_size < min_size:
        sr = net(torch.cat(lr_list, dim=0).contiguous())
        sr_list = list(sr.chunk(4, dim=0))
    else:
        sr_list = [forward_chop(net, patch.contiguous(), scale=scale, shave=shave, min_size=min_size) for patch in lr_list]
    h, w = scale * h, scale * w
    h_half, w_half = scale * h_half, scale * w_half
    h_size, w_size = scale * h_size, scale * w_size
    output = x.new_empty(b, c, h, w)
    output[:, :, 0:h_half, 0:w_half] = sr_list[0][:, :, 0:h_half, 0:w_half]
    output[:, :, 0:h_half, w_half:w] = sr_list[1][:, :, 0:h_half, (w_size - w + w_half):w_size]
    output[:, :, h_half:h, 0:w_half] = sr_list[2][:, :, (h_size - h + h_half):h_size, 0:w_half]
    output[:, :, h_half:h, w_half:w] = sr_list[3][:, :, (h_size - h + h_half):h_size, (w_size - w + w_half):w_size]
    return output
