"""Golden vectors for SURVEY §8(f) row 3 — the S2M network (model/s2m/s2m_network.py:55-65,
deeplabv3plus_resnet50) and S2MController.interact (interact/s2m_controller.py:22-37) — produced
by the UNMODIFIED reference modules imported from /root/reference (oracle/refshim.py), strict-
loaded with our seeded 368-tensor state dict, and checked against the oracle restatement on the
spot.  Run in the build container:  python -m oracle.gen_golden_s2m

TEST INFRASTRUCTURE (see oracle/stm_oracle.py)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim, s2m_oracle as S, weights as Wt  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def make_inputs(h=96, w=128, seed=21):
    """A 6-channel S2M input as the callers build it: normalised frame, hard previous mask,
    positive / negative scribble maps (davis_processor.py:57-66)."""
    g = torch.Generator().manual_seed(seed)
    image = torch.randn((1, 3, h, w), generator=g)
    prev = torch.zeros((1, 1, h, w))
    prev[:, :, h // 4:h // 2, w // 3:2 * w // 3] = 1
    scr = torch.zeros((1, 2, h, w))
    scr[:, 0, h // 3:h // 3 + 3, w // 4:w // 2] = 1
    scr[:, 1, 3 * h // 4:3 * h // 4 + 2, w // 8:w // 2] = 1
    return torch.cat([image, prev, scr], 1)


def make_controller_inputs(h=88, w=120, k=2, seed=22):
    """Unpadded scribble labels [h,w] (0 = background stroke, 255 = no stroke), padded image and
    previous label map as InferenceCore hands them to the controller (interactive_gui.py:836-838)."""
    g = torch.Generator().manual_seed(seed)
    nh, nw = (h + 15) // 16 * 16, (w + 15) // 16 * 16
    image = torch.randn((1, 3, nh, nw), generator=g)
    prev = torch.zeros((1, nh, nw), dtype=torch.int64)
    prev[:, nh // 5:nh // 2, nw // 6:nw // 2] = 1
    prev[:, nh // 2:3 * nh // 4, nw // 2:5 * nw // 6] = 2
    scr = np.full((h, w), 255, dtype=np.uint8)
    scr[h // 3:h // 3 + 2, w // 5:w // 2] = 1
    scr[2 * h // 3:2 * h // 3 + 2, w // 2:4 * w // 5] = 2
    scr[5:7, 5:w // 3] = 0
    return image, prev, scr, k


def main():
    torch.set_grad_enabled(False)
    assert refshim.available()
    sd = Wt.make_s2m_state_dict()
    report = {"torch": torch.__version__, "s2m_keys": len(sd)}
    with refshim.reference_on_path():
        from interact.s2m_controller import S2MController
        from model.s2m.s2m_network import deeplabv3plus_resnet50 as S2M

        net = S2M().eval()
        assert sorted(net.state_dict()) == sorted(sd)
        net.load_state_dict(sd, strict=True)

        x = make_inputs()
        feats = net.backbone(x)
        ref_logits = net(x)
        low, out = S.backbone(sd, x)
        ora_logits = S.s2m_forward(sd, x)
        ref_aspp = net.classifier.aspp(feats["out"])
        ora_aspp = S.aspp(sd, out)
        for name, a, b in (("s2m_low_level", feats["low_level"], low), ("s2m_layer4", feats["out"], out),
                           ("s2m_aspp", ref_aspp, ora_aspp), ("s2m_logits", ref_logits, ora_logits)):
            d = float((a - b).abs().max())
            report[name] = {"max_abs_diff_oracle_vs_reference": d, "ref_abs_max": float(a.abs().max())}
            assert d <= 1e-5 * max(1.0, float(a.abs().max())), (name, d)
        np.savez_compressed(os.path.join(OUT, "s2m_net.npz"), x=x.numpy(), logits=ref_logits.numpy(),
                            aspp=ref_aspp.numpy(), layer4=feats["out"].numpy().astype(np.float16),
                            low_level=feats["low_level"].numpy().astype(np.float16))

        image, prev, scr, k = make_controller_inputs()
        ctrl = S2MController(net, k, ignore_class=255, device="cpu")
        ref_m = ctrl.interact(image, prev, scr)
        ora_m = S.s2m_controller_interact(sd, image, prev, scr, k)
        d = float((ref_m - ora_m).abs().max())
        report["s2m_controller"] = {"max_abs_diff_oracle_vs_reference": d, "ref_abs_max": float(ref_m.abs().max()),
                                    "ref_min": float(ref_m.min())}
        assert d <= 1e-6, d
        np.savez_compressed(os.path.join(OUT, "s2m_controller.npz"), image=image.numpy(), prev=prev.numpy(), scr=scr,
                            k=np.int64(k), mask=ref_m.numpy())
    man = json.load(open(os.path.join(OUT, "MANIFEST.json")))
    man["s2m"] = report
    json.dump(man, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
