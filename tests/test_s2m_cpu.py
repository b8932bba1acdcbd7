"""CPU: SURVEY §8(f) row 3 — the S2M oracle against the reference-generated golden vectors, the
drop-in surface (state-dict keys, import paths, no CPU path) and the host-side input builder."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


@pytest.fixture(scope="module")
def s2m_sd():
    from oracle import weights
    return weights.make_s2m_state_dict()


def test_manifest_says_s2m_oracle_equals_reference():
    man = json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))["s2m"]
    assert man["s2m_keys"] == 368
    for k in ("s2m_low_level", "s2m_layer4", "s2m_aspp", "s2m_logits", "s2m_controller"):
        assert man[k]["max_abs_diff_oracle_vs_reference"] <= 1e-6, k


def test_s2m_oracle_matches_reference_golden(golden, s2m_sd):
    from oracle import s2m_oracle as S
    g = golden("s2m_net.npz")
    x = torch.from_numpy(g["x"])
    low, out = S.backbone(s2m_sd, x)
    assert float((low - torch.from_numpy(g["low_level"]).float()).abs().max()) <= 1e-2  # fixture stored as fp16
    assert float((out - torch.from_numpy(g["layer4"]).float()).abs().max()) <= 1e-2
    assert float((S.aspp(s2m_sd, out) - torch.from_numpy(g["aspp"])).abs().max()) <= 1e-5
    logits = S.s2m_forward(s2m_sd, x)
    assert logits.shape == (1, 1, 96, 128)
    assert float((logits - torch.from_numpy(g["logits"])).abs().max()) <= 1e-5
    assert 0.3 < float((logits > 0).float().mean()) < 0.7  # the fixture is not a trivial all-one mask


def test_s2m_controller_oracle_matches_reference_golden(golden, s2m_sd):
    from oracle import s2m_oracle as S
    g = golden("s2m_controller.npz")
    m = S.s2m_controller_interact(s2m_sd, torch.from_numpy(g["image"]), torch.from_numpy(g["prev"]), g["scr"], int(g["k"]))
    assert m.shape == (2, 1, 96, 128)  # 88x120 scribbles padded to multiples of 16
    assert float((m - torch.from_numpy(g["mask"])).abs().max()) <= 1e-6


def test_s2m_dropin_surface(s2m_sd):
    """Same import paths, constructor and checkpoint format as the reference (davis_processor.py:8,
    interactive_gui.py:30,34,1003-1004)."""
    from interact.s2m_controller import S2MController
    from model.s2m.s2m_network import deeplabv3plus_resnet50 as S2M
    import mivos_b200
    from mivos_b200 import arch, synth

    net = S2M()
    assert isinstance(net, mivos_b200.S2MNetwork) and not net.training
    assert list(net.state_dict()) == list(s2m_sd) and len(s2m_sd) == 368
    assert all(net.state_dict()[k].shape == v.shape for k, v in s2m_sd.items())
    net.load_state_dict(s2m_sd, strict=True)
    b = synth.make_s2m_state_dict()
    assert list(b) == list(s2m_sd) and all(torch.equal(b[k], s2m_sd[k]) for k in b)
    assert arch.S2M_ASPP_RATES == (6, 12, 18) and arch.S2M_DILATION[-1] == (1, 2) and arch.S2M_STRIDES[-1] == 1
    ctrl = S2MController(net, 2, ignore_class=255, device="cpu")
    assert ctrl.num_objects == 2 and ctrl.ignore_class == 255 and ctrl.s2m_net is net
    with pytest.raises(mivos_b200._lib.MivosError):
        net(torch.zeros((1, 6, 32, 32)))  # parameters on the CPU: there is no CPU path
    with pytest.raises(mivos_b200._lib.MivosError):
        net._prep(torch.zeros((1, 6, 30, 32)))  # callers pad to multiples of 16 first
    with pytest.raises(mivos_b200._lib.MivosError):
        S2M(num_classes=3)


def test_scribble_inputs_equal_the_reference_loop(golden):
    """The batched [K,6,nh,nw] input equals what s2m_controller.py:28-34 builds object by object."""
    from mivos_b200.s2m import scribble_inputs
    from oracle.stm_oracle import pad_divide_by
    g = golden("s2m_controller.npz")
    image, prev, scr, k = torch.from_numpy(g["image"]), torch.from_numpy(g["prev"]), g["scr"], int(g["k"])
    ids = np.arange(1, k + 1).reshape(-1, 1, 1)
    x = scribble_inputs(image, prev, scr[None] == ids, (scr[None] != ids) & (scr[None] != 255))
    for ki in range(1, k + 1):
        p_srb = (scr == ki).astype(np.uint8)
        n_srb = ((scr != ki) * (scr != 255)).astype(np.uint8)
        rs = torch.from_numpy(np.stack([p_srb, n_srb], 0)).unsqueeze(0).float()
        rs, _ = pad_divide_by(rs, 16, rs.shape[-2:])
        ref = torch.cat([image, (prev == ki).float().unsqueeze(0), rs], 1)
        assert torch.equal(x[ki - 1:ki], ref)


def test_s2m_packing_shapes(s2m_sd):
    """Packed weights of the dilated / pooled / concatenated layers have the K extents the gathers
    produce (host arithmetic only)."""
    from mivos_b200 import ops
    for dt, kq in ((torch.float16, 64), (torch.float32, 32)):
        pc = ops.pack_conv(s2m_sd["backbone.conv1.weight"], None, stride=2, im2col=True, device="cpu", dtype=dt)
        assert (pc.taps, pc.cin_pad) == (1, 320) and pc.cin_pad % kq == 0  # 49 x 6 = 294 -> 320
        pc = ops.pack_conv(s2m_sd["classifier.aspp.convs.2.0.weight"], None, im2col=True, device="cpu", dtype=dt)
        assert (pc.taps, pc.cin_pad, pc.cout_pad) == (1, 9 * 2048, 256)
        pc = ops.pack_conv(s2m_sd["classifier.classifier.0.weight"], None, device="cpu", dtype=dt)
        assert (pc.taps, pc.cin_pad) == (9, 320)  # 48 + 256 = 304 channels -> 320
        pc = ops.pack_conv(s2m_sd["classifier.project.0.weight"], None, device="cpu", dtype=dt)
        assert (pc.cout, pc.cout_pad) == (48, 64)
    w = s2m_sd["backbone.layer4.1.conv2.weight"]
    pc = ops.pack_conv(w, None, im2col=True, device="cpu", dtype=torch.float32)
    # k = (ky*3 + kx)*cin + ci: the order mivos_gather_dilated writes
    assert torch.allclose(pc.weight[0, 5, (1 * 3 + 2) * 512 + 7], ops.round_tf32(w[5, 7, 1, 2].reshape(1))[0])


def test_s2m_layer_graph_on_the_abi_emulator(golden, s2m_sd, monkeypatch):
    """The host side of engine.S2MEngine (packing, concat windows, dilations, gather orders, buffer
    shapes) driven over a PyTorch-CPU emulation of the C-ABI operators reproduces the reference's
    logits and intermediate features.  Tolerance: the packed weights are rounded to TF32."""
    import abi_emulator
    from mivos_b200 import engine, ops
    abi_emulator.install(monkeypatch, ops)
    g = golden("s2m_net.npz")
    x = torch.from_numpy(g["x"])
    eng = engine.S2MEngine(s2m_sd, "cpu", act_dtype=torch.float32)
    logits = eng.forward(x)
    ref = torch.from_numpy(g["logits"])
    assert logits.shape == ref.shape
    H, W = x.shape[-2:]
    low = eng.ws.halo("low", 1, H // 4, W // 4, 256)[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
    proj = eng.ws.halo("aspp_proj", 1, H // 16, W // 16, 256)[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
    for name, got, want in (("low_level", low, torch.from_numpy(g["low_level"]).float()),
                            ("aspp", proj, torch.from_numpy(g["aspp"])), ("logits", logits, ref)):
        err = float((got - want).abs().max()) / float(want.abs().max())
        assert err <= 5e-3, (name, err)
    # the zero border of every HALO buffer survived the whole pass
    for key, buf in eng.ws._bufs.items():
        if key[0] == "halo":
            assert float(buf[:, 0].abs().max()) == 0 and float(buf[:, -1].abs().max()) == 0, key
            assert float(buf[:, :, 0].abs().max()) == 0 and float(buf[:, :, -1].abs().max()) == 0, key
    # sigmoid variant and batch of two objects (rows of a batch are independent)
    x2 = torch.cat([x, x.flip(-1)], 0)
    p2 = eng.forward(x2, sigmoid=True)
    assert float((p2[0:1] - torch.sigmoid(ref)).abs().max()) <= 5e-3
