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


# ------------------------------------------------------------------ IoU helpers (util/tensor_util.py:5-60)
# Host-side metrics the DAVIS driver imports next to pad_divide_by (davis_processor.py:9); plain torch /
# numpy on whatever device the caller's masks live on.  intersection / union counts are float32 like the
# reference's, IoU = (I + 1e-6) / (U + 1e-6).
_IOU_EPS = 1e-6


def compute_tensor_iu(seg: torch.Tensor, gt: torch.Tensor):
    return (seg & gt).sum(dtype=torch.float32), (seg | gt).sum(dtype=torch.float32)


def compute_np_iu(seg, gt):
    import numpy as np
    return np.float32(np.count_nonzero(seg & gt)), np.float32(np.count_nonzero(seg | gt))


def compute_tensor_iou(seg: torch.Tensor, gt: torch.Tensor):
    inter, union = compute_tensor_iu(seg, gt)
    return (inter + _IOU_EPS) / (union + _IOU_EPS)


def compute_np_iou(seg, gt):
    inter, union = compute_np_iu(seg, gt)
    return (inter + _IOU_EPS) / (union + _IOU_EPS)


def _mean_iou(ious, count):
    # `count` keeps the type the reference divides by (a Python int, or numpy's gt.max()): it decides whether
    # the result is float32 or float64 under numpy's promotion rules
    return (sum(ious) + _IOU_EPS) / (count + _IOU_EPS)


def compute_multi_class_iou(seg: torch.Tensor, gt: torch.Tensor):
    """seg [K+1,h,w] scores incl. background, gt [K,1,h,w] soft masks: mean IoU of argmax(seg) == k+1 vs gt_k > 0.5."""
    pred = torch.argmax(seg, dim=0)
    return _mean_iou([compute_tensor_iou(pred == k + 1, gt[k, 0] > 0.5) for k in range(gt.shape[0])], gt.shape[0])


def compute_multi_class_iou_idx(seg, gt):
    """seg [h,w] label map, gt [K,h,w] soft masks (numpy)."""
    return _mean_iou([compute_np_iou(seg == k + 1, gt[k] > 0.5) for k in range(gt.shape[0])], gt.shape[0])


def compute_multi_class_iou_both_idx(seg, gt):
    """seg, gt [h,w] label maps (numpy); classes 1..gt.max()."""
    num_classes = gt.max()
    return _mean_iou([compute_np_iou(seg == k, gt == k) for k in range(1, int(num_classes) + 1)], num_classes)
