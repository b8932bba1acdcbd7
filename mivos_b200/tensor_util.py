"""Drop-in pad helpers (reference: util/tensor_util.py:62-93)."""
from __future__ import annotations

import torch

from . import ops


def pad_amounts(h: int, w: int, d: int):
    """(lw, uw, lh, uh) exactly as util/tensor_util.py:62-78 computes them."""
    nh = h if h % d == 0 else h + d - h % d
    nw = w if w % d == 0 else w + d - w % d
    lh, lw = (nh - h) // 2, (nw - w) // 2
    return (lw, nw - w - lw, lh, nh - h - lh)


def pad_divide_by(in_img: torch.Tensor, d: int, in_size=None):
    h, w = in_img.shape[-2:] if in_size is None else in_size
    pad = pad_amounts(h, w, d)
    if sum(pad) == 0:
        return in_img, pad
    if in_img.is_cuda and in_img.dtype == torch.float32:
        return ops.pad2d(in_img.contiguous(), pad), pad
    return torch.nn.functional.pad(in_img, pad), pad  # host-side staging of CPU clips


def unpad(img: torch.Tensor, pad):
    if pad[2] + pad[3] > 0:
        img = img[:, :, pad[2]:img.shape[2] - pad[3], :]
    if pad[0] + pad[1] > 0:
        img = img[:, :, :, pad[0]:img.shape[3] - pad[1]]
    return img


def unpad_3dim(img: torch.Tensor, pad):
    if pad[2] + pad[3] > 0:
        img = img[:, pad[2]:img.shape[1] - pad[3], :]
    if pad[0] + pad[1] > 0:
        img = img[:, :, pad[0]:img.shape[2] - pad[1]]
    return img
