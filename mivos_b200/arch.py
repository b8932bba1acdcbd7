"""Architecture tables of the propagation path and a parameter-tree builder.

The reference defines its networks as nn.Module classes (model/propagation/modules.py:38-114,
mod_resnet.py:114-150, prop_net.py:14-22,131-142, model/fusion_net.py:8-30).  Our kernels do not
run those modules; we only need (a) parameter containers whose ``state_dict`` keys and shapes are
identical to the reference's checkpoints (597 tensors for PropagationNetwork, 12
for FusionNet) and (b) the layer graph, which lives in ``engine.py``.  Both are generated from
the tables below.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch
import torch.nn as nn

# ("conv", name, cout, cin, ksize, has_bias) | ("bn", name, channels)
Entry = Tuple

TRUNK_PLANES = (64, 128, 256)
TRUNK_BLOCKS = (3, 4, 6)
TRUNK_STRIDES = (1, 2, 2)


def trunk_entries(prefix: str, cin: int, conv_bias: bool, layer_names) -> List[Entry]:
    out: List[Entry] = [("conv", f"{prefix}.conv1", 64, cin, 7, conv_bias), ("bn", f"{prefix}.bn1", 64)]
    inplanes = 64
    for lname, planes, blocks in zip(layer_names, TRUNK_PLANES, TRUNK_BLOCKS):
        for b in range(blocks):
            p = f"{prefix}.{lname}.{b}"
            out += [("conv", f"{p}.conv1", planes, inplanes, 1, conv_bias), ("bn", f"{p}.bn1", planes),
                    ("conv", f"{p}.conv2", planes, planes, 3, conv_bias), ("bn", f"{p}.bn2", planes),
                    ("conv", f"{p}.conv3", 4 * planes, planes, 1, conv_bias), ("bn", f"{p}.bn3", 4 * planes)]
            if b == 0:
                out += [("conv", f"{p}.downsample.0", 4 * planes, inplanes, 1, conv_bias),
                        ("bn", f"{p}.downsample.1", 4 * planes)]
            inplanes = 4 * planes
    return out


def resblock_entries(p: str, cin: int, cout: int) -> List[Entry]:
    out: List[Entry] = []
    if cin != cout:
        out.append(("conv", f"{p}.downsample", cout, cin, 3, True))
    out += [("conv", f"{p}.conv1", cout, cin, 3, True), ("conv", f"{p}.conv2", cout, cout, 3, True)]
    return out


def upblock_entries(p: str, skip_c: int, up_c: int, out_c: int) -> List[Entry]:
    return ([("conv", f"{p}.skip_conv1", up_c, skip_c, 3, True)] + resblock_entries(f"{p}.skip_conv2", up_c, up_c)
            + resblock_entries(f"{p}.out_conv", up_c, out_c))


MASK_LAYERS = ("layer1", "layer2", "layer3")  # MaskRGBEncoder attribute names (modules.py:49-51)
RGB_LAYERS = ("res2", "layer2", "layer3")     # RGBEncoder attribute names (modules.py:76-78)


def propagation_entries() -> List[Entry]:
    e = trunk_entries("mask_rgb_encoder", 5, True, MASK_LAYERS)   # biased convs: mod_resnet.py:81-86,119
    e += trunk_entries("rgb_encoder", 3, False, RGB_LAYERS)        # torchvision ResNet-50: no conv bias
    for kv in ("kv_m_f16", "kv_q_f16"):                            # prop_net.py:137-138
        e += [("conv", f"{kv}.key_proj", 128, 1024, 3, True), ("conv", f"{kv}.val_proj", 512, 1024, 3, True)]
    e += resblock_entries("decoder.compress", 1024, 512)           # prop_net.py:17-21
    e += upblock_entries("decoder.up_16_8", 512, 512, 256)
    e += upblock_entries("decoder.up_8_4", 256, 256, 256)
    e.append(("conv", "decoder.pred", 1, 256, 3, True))
    return e


def attention_read_entries() -> List[Entry]:
    """AttentionReadNetwork (model/attn_network.py:30-41): the two encoders and key/value heads of
    the propagation network, no decoder — same names, so a propagation checkpoint loads with
    strict=False (model/fusion_model.py:187)."""
    return [e for e in propagation_entries() if not e[1].startswith("decoder.")]


def fusion_entries() -> List[Entry]:
    return [("conv", "conv1.0", 32, 9, 3, True), ("conv", "conv2.0", 32, 32, 3, True),
            ("conv", "conv2.2", 32, 32, 3, True), ("conv", "conv3.0", 32, 32, 3, True),
            ("conv", "conv3.2", 32, 32, 3, True), ("conv", "final_conv", 1, 32, 3, True)]


# ---- Scribble-to-Mask network: DeepLabV3+ on a 6-channel ResNet-50, output stride 16
S2M_LAYERS = ("layer1", "layer2", "layer3", "layer4")
S2M_PLANES = (64, 128, 256, 512)
S2M_BLOCKS = (3, 4, 6, 3)
S2M_STRIDES = (1, 2, 2, 1)        # layer4: stride replaced by dilation (s2m_network.py:13-15, s2m_resnet.py:119-122)
S2M_DILATION = ((1, 1), (1, 1), (1, 1), (1, 2))  # (first block, remaining blocks) of each layer
S2M_ASPP_RATES = (6, 12, 18)      # s2m_network.py:15


def s2m_entries() -> List[Entry]:
    """Parameter table of model/s2m (s2m_resnet.py:70-148, _deeplab.py:30-58,119-160): 368 tensors,
    key names as the reference's DeepLabV3(backbone, classifier).state_dict() produces them."""
    e: List[Entry] = [("conv", "backbone.conv1", 64, 6, 7, False), ("bn", "backbone.bn1", 64)]
    cin = 64
    for lname, planes, blocks in zip(S2M_LAYERS, S2M_PLANES, S2M_BLOCKS):
        for b in range(blocks):
            p = f"backbone.{lname}.{b}"
            e += [("conv", f"{p}.conv1", planes, cin, 1, False), ("bn", f"{p}.bn1", planes),
                  ("conv", f"{p}.conv2", planes, planes, 3, False), ("bn", f"{p}.bn2", planes),
                  ("conv", f"{p}.conv3", 4 * planes, planes, 1, False), ("bn", f"{p}.bn3", 4 * planes)]
            if b == 0:
                e += [("conv", f"{p}.downsample.0", 4 * planes, cin, 1, False), ("bn", f"{p}.downsample.1", 4 * planes)]
            cin = 4 * planes
    c = "classifier"
    e += [("conv", f"{c}.project.0", 48, 256, 1, False), ("bn", f"{c}.project.1", 48),
          ("conv", f"{c}.aspp.convs.0.0", 256, 2048, 1, False), ("bn", f"{c}.aspp.convs.0.1", 256)]
    for i in range(1, 4):
        e += [("conv", f"{c}.aspp.convs.{i}.0", 256, 2048, 3, False), ("bn", f"{c}.aspp.convs.{i}.1", 256)]
    e += [("conv", f"{c}.aspp.convs.4.1", 256, 2048, 1, False), ("bn", f"{c}.aspp.convs.4.2", 256),
          ("conv", f"{c}.aspp.project.0", 256, 1280, 1, False), ("bn", f"{c}.aspp.project.1", 256),
          ("conv", f"{c}.classifier.0", 256, 304, 3, False), ("bn", f"{c}.classifier.1", 256),
          ("conv", f"{c}.classifier.3", 1, 256, 1, True)]
    return e


class ParamNode(nn.Module):
    """A bare container: children and parameters are attached by name so that state_dict keys
    reproduce the reference checkpoint format.  It has no forward — the kernels do the work."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("ParamNode only holds parameters; use the owning network's methods")


def _descend(root: nn.Module, dotted: str) -> nn.Module:
    node = root
    for part in dotted.split("."):
        nxt = node._modules.get(part)
        if nxt is None:
            nxt = ParamNode()
            node.add_module(part, nxt)
        node = nxt
    return node


def build_param_tree(root: nn.Module, entries: List[Entry], generator: torch.Generator) -> None:
    """Attach parameters/buffers for every entry.  Initial values are He-normal convs and identity-
    like BatchNorm statistics (the reference initialises its encoders from ImageNet weights fetched
    over the network — modules.py:42,70 — which is impossible offline; real use loads a checkpoint
    with load_state_dict exactly as eval_interactive_davis.py:58-68 does)."""
    for ent in entries:
        if ent[0] == "conv":
            _, name, cout, cin, ks, has_bias = ent
            node = _descend(root, name)
            std = math.sqrt(2.0 / (cin * ks * ks))
            node.register_parameter("weight", nn.Parameter(torch.randn((cout, cin, ks, ks), generator=generator) * std,
                                                           requires_grad=False))
            if has_bias:
                node.register_parameter("bias", nn.Parameter(torch.zeros(cout), requires_grad=False))
        else:
            _, name, c = ent
            node = _descend(root, name)
            node.register_parameter("weight", nn.Parameter(torch.ones(c), requires_grad=False))
            node.register_parameter("bias", nn.Parameter(torch.zeros(c), requires_grad=False))
            node.register_buffer("running_mean", torch.zeros(c))
            node.register_buffer("running_var", torch.ones(c))
            node.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
