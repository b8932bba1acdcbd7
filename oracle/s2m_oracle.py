"""CPU oracle for SURVEY.md §8(f) row 3 — the Scribble-to-Mask network (S2M, DeepLabV3+ on a
6-channel ResNet-50) and the two call sites that run it right before the propagation path.

TEST INFRASTRUCTURE ONLY (see oracle/stm_oracle.py's header): only ``tests/`` may import it.

Functional restatement in PyTorch-CPU fp32 driven by a reference-format ``state_dict`` (368
tensors, ``oracle/weights.py::s2m_spec``).  Pinned against the UNMODIFIED reference module by
``oracle/gen_golden_s2m.py`` (fixtures ``tests/golden/s2m_*.npz``).  Reference lines followed
(paths relative to the reference root):
  * backbone       model/s2m/s2m_resnet.py:27-66 (Bottleneck), :70-148 (ResNet, 6-channel conv1,
                   replace_stride_with_dilation=[False, False, True] for output stride 16 —
                   model/s2m/s2m_network.py:13-15,17-19)
  * head           model/s2m/_deeplab.py:30-58 (DeepLabHeadV3Plus), :119-160 (ASPPConv,
                   ASPPPooling, ASPP; Dropout is the identity in eval mode)
  * final resize   model/s2m/utils.py:16-21
  * call sites     interact/s2m_controller.py:22-37, davis_processor.py:55-68
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from .stm_oracle import _bn, pad_divide_by

SD = Dict[str, torch.Tensor]
ASPP_RATES = (6, 12, 18)  # s2m_network.py:15 (output_stride 16)


def _conv(sd: SD, name: str, x, stride=1, padding=0, dilation=1):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding, dilation=dilation)


def _bottleneck(sd: SD, p: str, x, stride: int, dilation: int):
    """s2m_resnet.py:46-66: 1x1-bn-relu, 3x3(stride, dilation, padding=dilation)-bn-relu, 1x1-bn,
    + (downsampled) identity, relu."""
    out = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x)))
    out = F.relu(_bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out, stride=stride, padding=dilation, dilation=dilation)))
    out = _bn(sd, p + ".bn3", _conv(sd, p + ".conv3", out))
    if (p + ".downsample.0.weight") in sd:
        x = _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x, stride=stride))
    return F.relu(out + x)


def backbone(sd: SD, x: torch.Tensor):
    """-> (low_level = layer1 output [N,256,H/4,W/4], out = layer4 output [N,2048,H/16,W/16]).
    _make_layer (s2m_resnet.py:116-139): with dilate=True the stride of layer4 becomes 1, its first
    block keeps the previous dilation (1) and the following blocks use dilation 2."""
    x = F.relu(_bn(sd, "backbone.bn1", _conv(sd, "backbone.conv1", x, stride=2, padding=3)))
    x = F.max_pool2d(x, 3, 2, 1)
    low = None
    for lname, blocks, stride, dil_first, dil_rest in (("layer1", 3, 1, 1, 1), ("layer2", 4, 2, 1, 1),
                                                       ("layer3", 6, 2, 1, 1), ("layer4", 3, 1, 1, 2)):
        for b in range(blocks):
            x = _bottleneck(sd, f"backbone.{lname}.{b}", x, stride if b == 0 else 1, dil_first if b == 0 else dil_rest)
        if lname == "layer1":
            low = x
    return low, x


def aspp(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """_deeplab.py:141-160."""
    p = "classifier.aspp"
    res = [F.relu(_bn(sd, p + ".convs.0.1", _conv(sd, p + ".convs.0.0", x)))]
    for i, r in enumerate(ASPP_RATES, start=1):
        res.append(F.relu(_bn(sd, f"{p}.convs.{i}.1", _conv(sd, f"{p}.convs.{i}.0", x, padding=r, dilation=r))))
    g = F.adaptive_avg_pool2d(x, 1)
    g = F.relu(_bn(sd, p + ".convs.4.2", _conv(sd, p + ".convs.4.1", g)))
    res.append(F.interpolate(g, size=x.shape[-2:], mode="bilinear", align_corners=False))
    return F.relu(_bn(sd, p + ".project.1", _conv(sd, p + ".project.0", torch.cat(res, 1))))


def s2m_forward(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """DeepLabV3.forward (utils.py:16-21) with the V3+ head (_deeplab.py:48-52): logits [N,1,H,W]."""
    low, out = backbone(sd, x)
    lowp = F.relu(_bn(sd, "classifier.project.1", _conv(sd, "classifier.project.0", low)))
    a = F.interpolate(aspp(sd, out), size=lowp.shape[2:], mode="bilinear", align_corners=False)
    y = F.relu(_bn(sd, "classifier.classifier.1", _conv(sd, "classifier.classifier.0", torch.cat([lowp, a], 1), padding=1)))
    y = _conv(sd, "classifier.classifier.3", y)
    return F.interpolate(y, size=x.shape[-2:], mode="bilinear", align_corners=False)


def s2m_controller_interact(sd: SD, image: torch.Tensor, prev_mask: torch.Tensor, scr_mask: np.ndarray,
                            num_objects: int, ignore_class: int = 255) -> torch.Tensor:
    """S2MController.interact (interact/s2m_controller.py:22-37): image [1,3,nh,nw] (padded),
    prev_mask [1,nh,nw] (argmax labels, padded), scr_mask np [h,w] labels -> [K,1,nh,nw]."""
    h, w = image.shape[-2:]
    out = torch.zeros((num_objects, 1, h, w), dtype=torch.float32)
    for ki in range(1, num_objects + 1):
        p_srb = (scr_mask == ki).astype(np.uint8)
        n_srb = ((scr_mask != ki) * (scr_mask != ignore_class)).astype(np.uint8)
        rs = torch.from_numpy(np.stack([p_srb, n_srb], 0)).unsqueeze(0).float()
        rs, _ = pad_divide_by(rs, 16, rs.shape[-2:])
        inputs = torch.cat([image, (prev_mask == ki).float().unsqueeze(0), rs], 1)
        out[ki - 1] = torch.sigmoid(s2m_forward(sd, inputs))
    return out
