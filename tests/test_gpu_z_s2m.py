"""GPU: SURVEY §8(f) row 3 — the S2M operators against plain PyTorch statements of the same ops
(tests/abi_emulator.py: F.unfold / F.interpolate / mean on the CPU), and the S2M network and
controller against the vectors the UNMODIFIED reference produced (tests/golden/s2m_*.npz).
Gathers and the pooled broadcast of fp16-representable values are bit-exact; bilinear resizes equal
ATen's to fp32 rounding (1e-6) plus one fp16 rounding for fp16 maps; network outputs carry the conv
tolerances of DESIGN.md §6 (fp16 / TF32 operands, fp32 accumulation)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import abi_emulator as E  # noqa: E402
from mivos_b200 import _lib, ops  # noqa: E402

DT = {"fp16": torch.float16, "tf32": torch.float32}


def to_halo(x, cpad=None, dtype=torch.float32):
    n, c, h, w = x.shape
    hb = torch.zeros((n, h + 2, w + 2, cpad or c), device=x.device, dtype=dtype)
    hb[:, 1:-1, 1:-1, :c] = x.permute(0, 2, 3, 1).to(dtype)
    return hb


def _sync():
    torch.cuda.synchronize()
    _lib.poll_kernel_error()


def _rand16(shape, seed):
    """fp32 values that are exactly representable in fp16 (so copies are exact in both map types)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).half().float()


@pytest.mark.parametrize("act", ["fp16", "tf32"])
@pytest.mark.parametrize("n,h,w,c,dil", [(1, 6, 8, 64, 2), (2, 30, 54, 128, 6), (1, 30, 54, 64, 18), (1, 9, 7, 32, 12)])
def test_gather_dilated(dev, act, n, h, w, c, dil):
    dt = DT[act]
    x = _rand16((n, c, h, w), 3 * c + dil)
    xh_cpu = to_halo(x)
    rows = n * (h + 2) * (w + 2)
    want = E.gather_dilated(xh_cpu, n, h, w, c, dil, torch.full((rows, 9 * c), 5.0))
    got = torch.full((rows, 9 * c), 5.0, device=dev, dtype=dt)
    ops.gather_dilated(to_halo(x.to(dev), dtype=dt), n, h, w, c, dil, got)
    _sync()
    assert torch.equal(got.float().cpu(), want)


@pytest.mark.parametrize("act", ["fp16", "tf32"])
def test_stem_gather_six_channels(dev, act):
    dt = DT[act]
    n, h, w = 2, 32, 48
    x = _rand16((n, 6, h, w), 17)
    rows = n * (h // 2 + 2) * (w // 2 + 2)
    want = E.stem_gather_frames(x, torch.zeros((rows, 320)))
    got = torch.full((rows, 320), 3.0, device=dev, dtype=dt)
    ops.stem_gather_frames(x.to(dev), got)
    _sync()
    assert torch.equal(got.float().cpu(), want)
    # the 3-channel entry is the batch form of mivos_stem_gather
    x3 = _rand16((n, 3, h, w), 18)
    kp = 192 if act == "fp16" else 160
    a = torch.zeros((rows, kp), device=dev, dtype=dt)
    b = torch.zeros((rows, kp), device=dev, dtype=dt)
    ops.stem_gather_frames(x3.to(dev), a)
    ops.stem_gather(x3.to(dev), None, b)
    _sync()
    assert torch.equal(a, b)


@pytest.mark.parametrize("act", ["fp16", "tf32"])
@pytest.mark.parametrize("n,h,w,c", [(1, 6, 8, 2048), (3, 30, 54, 256), (1, 5, 3, 40)])
def test_avgpool_broadcast(dev, act, n, h, w, c):
    dt = DT[act]
    x = _rand16((n, c, h, w), c + h)
    cs = c + 24
    want = E.halo_avgpool_broadcast(to_halo(x, cs), n, h, w, c, torch.zeros((n, h + 2, w + 2, cs + 8)), in_coff=0, out_coff=8)
    src = torch.zeros((n, h + 2, w + 2, cs), device=dev, dtype=dt)
    src[:, 1:-1, 1:-1, 16:16 + c] = x.permute(0, 2, 3, 1).to(dev)
    got = torch.zeros((n, h + 2, w + 2, cs + 8), device=dev, dtype=dt)
    ops.halo_avgpool_broadcast(src, n, h, w, c, got, in_coff=16, out_coff=8)
    _sync()
    tol = 1e-3 if act == "fp16" else 1e-6  # fp32 sum in another order (+ one fp16 rounding of the mean)
    assert float((got.float().cpu() - want).abs().max()) <= tol
    assert float(got[:, 0].abs().max()) == 0 and float(got[:, :, -1].abs().max()) == 0  # border untouched
    assert float(got[..., :8].abs().max()) == 0 and float(got[..., 8 + c:].abs().max()) == 0  # window only


@pytest.mark.parametrize("act", ["fp16", "tf32"])
@pytest.mark.parametrize("n,hs,ws,h,w,c", [(1, 6, 8, 24, 32, 256), (2, 30, 54, 120, 216, 64), (1, 1, 1, 6, 8, 32),
                                           (1, 7, 5, 10, 13, 16)])
def test_upsample_bilinear_window(dev, act, n, hs, ws, h, w, c):
    dt = DT[act]
    x = _rand16((n, c, hs, ws), hs * 5 + c)
    want = E.upsample_bilinear(to_halo(x), n, hs, ws, torch.zeros((n, h + 2, w + 2, c + 64)), h, w, c, dst_coff=48)
    got = torch.zeros((n, h + 2, w + 2, c + 64), device=dev, dtype=dt)
    ops.upsample_bilinear(to_halo(x.to(dev), dtype=dt), n, hs, ws, got, h, w, c, dst_coff=48)
    _sync()
    tol = 4e-3 if act == "fp16" else 2e-6  # values up to ~4: half an fp16 ulp is 2e-3
    assert float((got.float().cpu() - want).abs().max()) <= tol
    assert float(got[:, 0].abs().max()) == 0 and float(got[:, :, 0].abs().max()) == 0
    assert float(got[..., :48].abs().max()) == 0 and float(got[..., 48 + c:].abs().max()) == 0


@pytest.mark.parametrize("sig", [False, True])
@pytest.mark.parametrize("n,hs,ws,H,W", [(1, 24, 32, 96, 128), (3, 120, 216, 480, 864), (2, 5, 7, 13, 9)])
def test_upsample_to_plane(dev, sig, n, hs, ws, H, W):
    x = torch.randn((n, 1, hs, ws), generator=torch.Generator().manual_seed(hs)) * 3
    hb = torch.zeros((n, hs + 2, ws + 2, 32))
    hb[:, 1:-1, 1:-1, 4:5] = x.permute(0, 2, 3, 1)
    want = E.halo_upsample_to_plane(hb, n, hs, ws, H, W, coff=4, sigmoid=sig)
    got = ops.halo_upsample_to_plane(hb.to(dev), n, hs, ws, H, W, coff=4, sigmoid=sig)
    _sync()
    assert got.shape == (n, 1, H, W)
    assert float((got.cpu() - want).abs().max()) <= 1e-5


@pytest.mark.parametrize("act", ["fp16", "tf32"])
def test_conv_into_channel_window_with_ragged_cout(dev, act):
    """classifier.project: 1x1 256 -> 48 (+ReLU) written into channels [0, 48) of the 320-channel head
    buffer; channels 48.. must stay untouched."""
    dt = DT[act]
    g = torch.Generator().manual_seed(5)
    n, h, w = 2, 24, 32
    x = torch.randn((n, 256, h, w), generator=g)
    wt = torch.randn((48, 256, 1, 1), generator=g) / 16
    b = torch.randn((48,), generator=g)
    pc = ops.pack_conv(wt, b, device=dev, dtype=dt)
    assert (pc.cout, pc.cout_pad) == (48, 64)
    out = torch.full((n, h + 2, w + 2, 320), 7.0, device=dev, dtype=dt)
    ops.conv_gemm(to_halo(x.to(dev), dtype=dt), pc, n, h, w, out, relu=True)
    _sync()
    xr = x.half().float() if act == "fp16" else x
    want = torch.nn.functional.conv2d(xr.double(), pc.weight[0, :48, :256].double().cpu().view(48, 256, 1, 1), b.double()).relu()
    got = out[:, 1:-1, 1:-1, :48].permute(0, 3, 1, 2).double().cpu()
    assert float((got - want).abs().max()) <= 4e-3 * float(want.abs().max())
    assert bool((out[..., 48:] == 7).all()) and bool((out[:, 0] == 7).all())


# ------------------------------------------------------------------------------------------ network
@pytest.fixture(scope="module", params=["fp16", "tf32"])
def s2m_net(request, dev):
    import mivos_b200
    from oracle import weights
    net = mivos_b200.S2MNetwork(act_dtype=DT[request.param])
    net.load_state_dict(weights.make_s2m_state_dict(), strict=True)
    return net.to(dev)


def test_s2m_network_matches_reference_golden(golden, s2m_net, dev):
    g = golden("s2m_net.npz")
    x = torch.from_numpy(g["x"]).to(dev)
    before = int(_lib.load().mivos_launch_count())
    logits = s2m_net(x)
    _sync()
    assert int(_lib.load().mivos_launch_count()) - before > 60  # the whole pass ran through the C ABI
    ref = torch.from_numpy(g["logits"])
    assert logits.shape == ref.shape and logits.dtype == torch.float32
    eng = s2m_net.engine()
    H, W = x.shape[-2:]
    low = eng.ws.halo("low", 1, H // 4, W // 4, 256)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float().cpu()
    proj = eng.ws.halo("aspp_proj", 1, H // 16, W // 16, 256)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float().cpu()
    for name, got, want, tol in (("low_level", low, torch.from_numpy(g["low_level"]).float(), 1e-2),
                                 ("aspp", proj, torch.from_numpy(g["aspp"]), 1e-2),
                                 ("logits", logits.cpu(), ref, 2e-2)):
        err = float((got - want).abs().max()) / float(want.abs().max())
        print(f"s2m {name}: rel err {err:.2e}")
        assert err <= tol, (name, err)
    prob = s2m_net.forward_sigmoid(x).cpu()
    _sync()
    assert float((prob - torch.sigmoid(ref)).abs().max()) <= 3e-2
    assert float(((prob > 0.5) != (ref > 0)).float().mean()) <= 1e-2
    # a batch is the per-image result (objects of one interaction are batched); tile widths / split-K
    # factors depend on the row count, so the accumulation order — not the value — may differ
    p2 = s2m_net.forward_sigmoid(torch.cat([x.flip(-1), x], 0))
    _sync()
    assert float((p2[1:2].cpu() - prob).abs().max()) <= 5e-3


def test_s2m_controller_matches_reference_golden(golden, s2m_net, dev):
    from interact.s2m_controller import S2MController
    g = golden("s2m_controller.npz")
    ctrl = S2MController(s2m_net, int(g["k"]), ignore_class=255, device=dev)
    m = ctrl.interact(torch.from_numpy(g["image"]), torch.from_numpy(g["prev"]), g["scr"])
    _sync()
    ref = torch.from_numpy(g["mask"])
    assert m.shape == ref.shape and m.is_cuda
    assert float((m.cpu() - ref).abs().max()) <= 3e-2
    assert float(((m.cpu() > 0.5) != (ref > 0.5)).float().mean()) <= 1e-2


def test_s2m_feeds_the_propagation_path(nets, s2m_net, dev):
    """davis_processor.py:52-82: S2M probabilities -> aggregate_wbg(hard) -> InferenceCore.interact."""
    import mivos_b200
    from mivos_b200 import synth
    images, _ = synth.synthetic_clip(4, 96, 128, 2, seed=5)
    core = mivos_b200.InferenceCore(nets[20], None, images, 2, mem_freq=2, device=str(dev))
    scr = np.full((96, 128), 255, dtype=np.uint8)
    scr[30:33, 20:70] = 1
    scr[60:63, 50:110] = 2
    ctrl = mivos_b200.S2MController(s2m_net, 2, ignore_class=255, device=dev)
    prob = ctrl.interact(core.get_image_buffered(1), core.masks[1], scr)
    mask = mivos_b200.aggregate_wbg(prob, keep_bg=True, hard=True)
    out = core.interact(mask, 1)
    _sync()
    assert out.shape == (4, 96, 128) and out.dtype == np.uint8 and int(out.max()) <= 2
