"""Golden vectors for SURVEY §8(f) row 2 — `AttentionReadNetwork.forward` (model/attn_network.py:48-80)
— produced by the UNMODIFIED reference module imported from /root/reference (oracle/refshim.py),
loaded with our seeded propagation state dict exactly as the reference does
(`load_state_dict(prop_sd, strict=False)`, model/fusion_model.py:187), and checked against the
oracle restatement on the spot.  Run in the build container:  python -m oracle.gen_golden_attn

TEST INFRASTRUCTURE (see oracle/stm_oracle.py)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim, stm_oracle as O, weights as Wt  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def make_inputs(b=2, h=64, w=96, seed=11):
    g = torch.Generator().manual_seed(seed)
    image = torch.randn((b, 3, h, w), generator=g)
    query = torch.randn((b, 3, h, w), generator=g)
    masks = [torch.rand((b, 1, h, w), generator=g) for _ in range(4)]  # soft masks m11, m21, m12, m22
    return image, masks, query


def main():
    torch.set_grad_enabled(False)
    ref = refshim.load_reference()
    psd = Wt.make_prop_state_dict(1234)
    net = ref.build_attn(psd)
    image, (m11, m21, m12, m22), query = make_inputs()
    a1, a2 = net(image, m11, m21, m12, m22, query)
    o1, o2 = O.attention_read_network(psd, image, m11, m21, m12, m22, query)
    d = max(float((a1 - o1).abs().max()), float((a2 - o2).abs().max()))
    report = {"max_abs_diff_oracle_vs_reference": d, "ref_abs_max": float(max(a1.abs().max(), a2.abs().max())),
              "torch": torch.__version__}
    assert d <= 1e-6, d
    np.savez_compressed(os.path.join(OUT, "attn_read.npz"), image=image.numpy(), query=query.numpy(), m11=m11.numpy(),
                        m21=m21.numpy(), m12=m12.numpy(), m22=m22.numpy(), attn1=a1.numpy(), attn2=a2.numpy())
    ref._restore()
    man = json.load(open(os.path.join(OUT, "MANIFEST.json")))
    man["attn_read"] = report
    json.dump(man, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)
    print(json.dumps(report))


if __name__ == "__main__":
    main()
