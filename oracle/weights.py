"""Seeded synthetic weights in the reference's checkpoint format (TEST INFRASTRUCTURE, see
stm_oracle.py's header).

The trained checkpoints are Google-Drive downloads (download_model.py:8-14) and are unavailable
offline, so parity is established on seeded random weights whose KEY SET AND SHAPES equal the
reference's ``PropagationNetwork().state_dict()`` (597 tensors) and ``FusionNet().state_dict()``
(12 tensors); ``oracle/gen_golden.py`` proves that with a strict ``load_state_dict`` into the real
reference modules.  BatchNorm running statistics are randomised so that BN folding is exercised.
The table below is our own statement of the architecture (modules.py:38-114, mod_resnet.py:114-150,
prop_net.py:14-22,131-142, fusion_net.py:8-30), not a copy of the module code.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

# (name, cout, cin, ksize, has_bias) for convs; ("bn", name, channels) for BatchNorm
Entry = Tuple


def _resnet_trunk(prefix: str, cin: int, conv_bias: bool, layer_names: Tuple[str, str, str]) -> List[Entry]:
    e: List[Entry] = [("conv", f"{prefix}.conv1", 64, cin, 7, conv_bias), ("bn", f"{prefix}.bn1", 64)]
    inplanes = 64
    for lname, planes, blocks in zip(layer_names, (64, 128, 256), (3, 4, 6)):
        for b in range(blocks):
            p = f"{prefix}.{lname}.{b}"
            e += [("conv", p + ".conv1", planes, inplanes, 1, conv_bias), ("bn", p + ".bn1", planes),
                  ("conv", p + ".conv2", planes, planes, 3, conv_bias), ("bn", p + ".bn2", planes),
                  ("conv", p + ".conv3", planes * 4, planes, 1, conv_bias), ("bn", p + ".bn3", planes * 4)]
            if b == 0:
                e += [("conv", p + ".downsample.0", planes * 4, inplanes, 1, conv_bias),
                      ("bn", p + ".downsample.1", planes * 4)]
            inplanes = planes * 4
    return e


def _resblock(p: str, cin: int, cout: int) -> List[Entry]:
    e: List[Entry] = []
    if cin != cout:
        e.append(("conv", p + ".downsample", cout, cin, 3, True))
    e += [("conv", p + ".conv1", cout, cin, 3, True), ("conv", p + ".conv2", cout, cout, 3, True)]
    return e


def _upblock(p: str, skip_c: int, up_c: int, out_c: int) -> List[Entry]:
    return ([("conv", p + ".skip_conv1", up_c, skip_c, 3, True)] + _resblock(p + ".skip_conv2", up_c, up_c)
            + _resblock(p + ".out_conv", up_c, out_c))


def prop_spec() -> List[Entry]:
    e = _resnet_trunk("mask_rgb_encoder", 5, True, ("layer1", "layer2", "layer3"))
    e += _resnet_trunk("rgb_encoder", 3, False, ("res2", "layer2", "layer3"))
    for kv in ("kv_m_f16", "kv_q_f16"):
        e += [("conv", kv + ".key_proj", 128, 1024, 3, True), ("conv", kv + ".val_proj", 512, 1024, 3, True)]
    e += _resblock("decoder.compress", 1024, 512)
    e += _upblock("decoder.up_16_8", 512, 512, 256)
    e += _upblock("decoder.up_8_4", 256, 256, 256)
    e.append(("conv", "decoder.pred", 1, 256, 3, True))
    return e


def fusion_spec() -> List[Entry]:
    return [("conv", "conv1.0", 32, 9, 3, True), ("conv", "conv2.0", 32, 32, 3, True),
            ("conv", "conv2.2", 32, 32, 3, True), ("conv", "conv3.0", 32, 32, 3, True),
            ("conv", "conv3.2", 32, 32, 3, True), ("conv", "final_conv", 1, 32, 3, True)]


def s2m_spec() -> List[Entry]:
    """Scribble-to-Mask network = DeepLabV3+ / ResNet-50, 6 input channels, output stride 16
    (model/s2m/s2m_network.py:8-33, s2m_resnet.py:70-148, _deeplab.py:30-58,119-160): backbone
    conv1..layer4 (bias-free convs + BN), classifier.project (1x1 256->48), classifier.aspp
    (1x1, three dilated 3x3, pooled 1x1, 1x1 projection of the 1280-channel concat) and
    classifier.classifier (3x3 304->256 + BN, biased 1x1 256->1)."""
    e: List[Entry] = [("conv", "backbone.conv1", 64, 6, 7, False), ("bn", "backbone.bn1", 64)]
    inplanes = 64
    for lname, planes, blocks in zip(("layer1", "layer2", "layer3", "layer4"), (64, 128, 256, 512), (3, 4, 6, 3)):
        for b in range(blocks):
            p = f"backbone.{lname}.{b}"
            e += [("conv", p + ".conv1", planes, inplanes, 1, False), ("bn", p + ".bn1", planes),
                  ("conv", p + ".conv2", planes, planes, 3, False), ("bn", p + ".bn2", planes),
                  ("conv", p + ".conv3", planes * 4, planes, 1, False), ("bn", p + ".bn3", planes * 4)]
            if b == 0:
                e += [("conv", p + ".downsample.0", planes * 4, inplanes, 1, False),
                      ("bn", p + ".downsample.1", planes * 4)]
            inplanes = planes * 4
    e += [("conv", "classifier.project.0", 48, 256, 1, False), ("bn", "classifier.project.1", 48)]
    e += [("conv", "classifier.aspp.convs.0.0", 256, 2048, 1, False), ("bn", "classifier.aspp.convs.0.1", 256)]
    for i in (1, 2, 3):
        e += [("conv", f"classifier.aspp.convs.{i}.0", 256, 2048, 3, False), ("bn", f"classifier.aspp.convs.{i}.1", 256)]
    e += [("conv", "classifier.aspp.convs.4.1", 256, 2048, 1, False), ("bn", "classifier.aspp.convs.4.2", 256)]
    e += [("conv", "classifier.aspp.project.0", 256, 1280, 1, False), ("bn", "classifier.aspp.project.1", 256)]
    e += [("conv", "classifier.classifier.0", 256, 304, 3, False), ("bn", "classifier.classifier.1", 256),
          ("conv", "classifier.classifier.3", 1, 256, 1, True)]
    return e


def _fill(spec: List[Entry], seed: int, gain: Dict[str, float]) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for ent in spec:
        if ent[0] == "conv":
            _, name, cout, cin, ks, bias = ent
            std = math.sqrt(2.0 / (cin * ks * ks))
            for key, mult in gain.items():
                if key in name:
                    std *= mult
            sd[name + ".weight"] = torch.randn((cout, cin, ks, ks), generator=g) * std
            if bias:
                sd[name + ".bias"] = torch.randn((cout,), generator=g) * 0.05
        else:
            _, name, c = ent
            last = name.endswith("bn3") or name.endswith("downsample.1")
            lo, hi = (0.25, 0.55) if last else (0.7, 1.3)
            sd[name + ".weight"] = torch.rand((c,), generator=g) * (hi - lo) + lo
            sd[name + ".bias"] = torch.randn((c,), generator=g) * 0.1
            sd[name + ".running_mean"] = torch.randn((c,), generator=g) * 0.2
            sd[name + ".running_var"] = torch.rand((c,), generator=g) * 1.0 + 0.5
            sd[name + ".num_batches_tracked"] = torch.tensor(100, dtype=torch.long)
    return sd


def make_prop_state_dict(seed: int = 1234) -> "OrderedDict[str, torch.Tensor]":
    # gains keep activations O(1) through 50 un-normalised layers and logits in a useful range
    gain = {"decoder": 0.75, "key_proj": 1.0, "val_proj": 1.0, "decoder.pred": 4.0}
    return _fill(prop_spec(), seed, gain)


def make_fusion_state_dict(seed: int = 4321) -> "OrderedDict[str, torch.Tensor]":
    sd = _fill(fusion_spec(), seed, {"final_conv": 1.0})
    # random 3x3 stacks give a strongly negative logit; centre it so fused masks are non-trivial
    sd["final_conv.bias"] = sd["final_conv.bias"] + 2.6
    return sd


def make_s2m_state_dict(seed: int = 2468) -> "OrderedDict[str, torch.Tensor]":
    # the last 1x1 gets a gain so that the sigmoid output is not pinned near 0.5
    sd = _fill(s2m_spec(), seed, {"classifier.classifier.3": 6.0})
    # random stacks give an all-positive logit map; centre it so that masks are non-trivial
    sd["classifier.classifier.3.bias"] = sd["classifier.classifier.3.bias"] - 5.2
    return sd


def synthetic_clip(t: int, h: int, w: int, k: int, seed: int = 1234):
    """DAVIS-shaped synthetic clip: normalised-looking frames with temporal
    coherence (a smooth random field drifting over time, so that propagation has something to
    follow) and K disjoint rectangles as the one-hot first-frame mask (background included)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn((1, 3, h // 8 + 4, w // 8 + 4 + t), generator=g)
    frames = []
    for i in range(t):
        crop = base[:, :, :, i:i + w // 8 + 4]
        up = torch.nn.functional.interpolate(crop, size=(h + 32, w + 32), mode="bilinear", align_corners=False)
        frames.append(up[:, :, 16:16 + h, 16:16 + w])
    images = torch.stack(frames, 1) + 0.1 * torch.randn((1, t, 3, h, w), generator=g)
    mask = torch.zeros((k + 1, 1, h, w))
    for j in range(k):
        y0 = int(h * (0.15 + 0.6 * j / max(k, 1)))
        x0 = int(w * (0.1 + 0.7 * j / max(k, 1)))
        mask[j + 1, 0, y0:y0 + h // 4, x0:x0 + w // 5] = 1
    mask[0] = 1 - mask[1:].sum(0).clamp(0, 1)
    return images.contiguous(), mask
