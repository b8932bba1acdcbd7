"""Synthetic DAVIS-shaped clips and seeded checkpoint-format weights (there is no network for the
real DAVIS frames or the Google-Drive checkpoints — download_model.py:8-14).  Used by bench.py and
smoke(); the recipe is deliberately identical to the test oracle's generator (oracle/weights.py)
so both sides can be driven with the same numbers — tests/test_host_logic.py checks equality."""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from . import arch


def _fill(entries, seed: int, gain):
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for ent in entries:
        if ent[0] == "conv":
            _, name, cout, cin, ks, has_bias = ent
            std = math.sqrt(2.0 / (cin * ks * ks))
            for key, mult in gain.items():
                if key in name:
                    std *= mult
            sd[name + ".weight"] = torch.randn((cout, cin, ks, ks), generator=g) * std
            if has_bias:
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


def make_prop_state_dict(seed: int = 1234):
    return _fill(arch.propagation_entries(), seed, {"decoder": 0.75, "key_proj": 1.0, "val_proj": 1.0, "decoder.pred": 4.0})


def make_fusion_state_dict(seed: int = 4321):
    sd = _fill(arch.fusion_entries(), seed, {"final_conv": 1.0})
    sd["final_conv.bias"] = sd["final_conv.bias"] + 2.6
    return sd


def make_s2m_state_dict(seed: int = 2468):
    sd = _fill(arch.s2m_entries(), seed, {"classifier.classifier.3": 6.0})
    sd["classifier.classifier.3.bias"] = sd["classifier.classifier.3.bias"] - 5.2
    return sd


def synthetic_clip(t: int, h: int, w: int, k: int, seed: int = 1234):
    """[1,t,3,h,w] normalised-looking frames (a smooth random field drifting over time + noise)
    and a one-hot first-frame mask [(k+1),1,h,w] of k disjoint rectangles."""
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


def second_interaction_mask(k: int, h: int, w: int, seed: int = 77):
    """A different one-hot mask [(k+1),1,h,w] for a SECOND interaction on the same clip (cfg-4: the
    difference masks of fuse_one_frame, inference_core.py:233-235, must be non-trivial): the rectangles
    of synthetic_clip shifted by a seeded offset."""
    g = torch.Generator().manual_seed(seed)
    dy, dx = int(torch.randint(10, 40, (1,), generator=g)), int(torch.randint(10, 60, (1,), generator=g))
    mask = torch.zeros((k + 1, 1, h, w))
    for j in range(k):
        y0 = int(h * (0.15 + 0.6 * j / max(k, 1))) + dy
        x0 = int(w * (0.1 + 0.7 * j / max(k, 1))) + dx
        mask[j + 1, 0, y0:y0 + h // 4, x0:x0 + w // 5] = 1
    mask[0] = 1 - mask[1:].sum(0).clamp(0, 1)
    return mask
