"""GPU: the operator forms added in round 2.
  * space-to-depth stems (mivos_stem_gather_s2d + four vertical conv taps) vs the 7x7/stride-2 convolution, and
    bit-exact vs the header contract (tests/abi_emulator.py), frames / frame + masks / groups, fp32 and fp16;
  * batched forms for lock-step clips: stem groups, upsample2x_add skip_n, upsample4x_sigmoid_aggregate groups,
    halo_copy source broadcast — each equal (bit for bit) to the per-clip launches it replaces;
  * memory read with query sets (q_div): ONE call for C clips == C calls, bit for bit, on both generators.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import abi_emulator as E  # noqa: E402  (header contracts in PyTorch: checker only)
from mivos_b200 import _lib, ops  # noqa: E402


def _halo(x, dtype=torch.float32):
    n, c, h, w = x.shape
    out = torch.zeros((n, h + 2, w + 2, c), device=x.device, dtype=dtype)
    out[:, 1:-1, 1:-1] = x.permute(0, 2, 3, 1).to(dtype)
    return out


def _from_halo(hb, c):
    return hb[:, 1:-1, 1:-1, :c].permute(0, 3, 1, 2).float()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_s2d_stem_equals_the_7x7_stride2_convolution(dev, dtype):
    g = torch.Generator().manual_seed(11)
    H, W, K, C = 96, 160, 3, 2
    frames = torch.randn((C, 3, H, W), generator=g).to(dev)
    masks = torch.rand((C, K, 1, H, W), generator=g).to(dev)
    w5 = (torch.randn((64, 5, 7, 7), generator=g) / 15).to(dev)
    b = torch.randn((64,), generator=g).to(dev)
    rows = (H // 2 + 2) * (W // 2 + 2)
    tol = 2e-3
    # (a) batch of 3-channel frames
    pc3 = ops.pack_conv(w5[:, :3].contiguous(), b, stride=2, device=dev, dtype=dtype, stem_s2d=True)
    assert pc3.taps == 4 and pc3.cin_pad == (64 if dtype == torch.float16 else 64)
    g3 = torch.full((C * rows, pc3.cin_pad), 9.0, device=dev, dtype=dtype)
    ops.stem_gather(frames, None, g3, s2d=True)
    want = E.stem_gather(frames.cpu(), None, torch.zeros((C * rows, pc3.cin_pad)), s2d=True)
    assert torch.equal(g3.float().cpu(), want.to(dtype).float())  # bit-exact gather (one rounding to the element type)
    o = torch.zeros((C, H // 2 + 2, W // 2 + 2, 64), device=dev, dtype=dtype)
    ops.conv_gemm(g3, pc3, C, H // 2, W // 2, o, relu=True)
    xr = frames.to(dtype).double()
    ref = F.conv2d(xr, pc3_weight_as_conv(pc3, 3), b.double(), stride=2, padding=3).relu()
    assert float((_from_halo(o, 64).double() - ref).abs().max()) <= tol * float(ref.abs().max())
    # (b) one frame + K masks, and (c) the same as G groups in one launch
    pc5 = ops.pack_conv(w5, b, stride=2, device=dev, dtype=dtype, stem_s2d=True)
    per_clip = []
    for c in range(C):
        gm = torch.zeros((K * rows, pc5.cin_pad), device=dev, dtype=dtype)
        ops.stem_gather(frames[c:c + 1], masks[c], gm, s2d=True)
        per_clip.append(gm)
        others = masks[c].sum(0, keepdim=True) - masks[c]
        inp = torch.cat([frames[c:c + 1].expand(K, -1, -1, -1), masks[c], others], 1)
        o5 = torch.zeros((K, H // 2 + 2, W // 2 + 2, 64), device=dev, dtype=dtype)
        ops.conv_gemm(gm, pc5, K, H // 2, W // 2, o5)
        ref5 = F.conv2d(inp.to(dtype).double(), pc3_weight_as_conv(pc5, 5), b.double(), stride=2, padding=3)
        assert float((_from_halo(o5, 64).double() - ref5).abs().max()) <= tol * float(ref5.abs().max())
    # groups: masks as a strided view [C, K, 1, H, W] of a [C, K+1, ...] probability volume, like the lock-step step
    vol = torch.zeros((C, K + 1, 1, H, W), device=dev)
    vol[:, 1:] = masks
    gg = torch.zeros((C * K * rows, pc5.cin_pad), device=dev, dtype=dtype)
    ops.stem_gather(frames, vol[:, 1:], gg, s2d=True)
    assert torch.equal(gg, torch.cat(per_clip, 0))
    # the 49-tap im2col form takes groups too
    pci = ops.pack_conv(w5, b, stride=2, im2col=True, device=dev, dtype=dtype)
    gi = torch.zeros((C * K * rows, pci.cin_pad), device=dev, dtype=dtype)
    ops.stem_gather(frames, vol[:, 1:], gi)
    for c in range(C):
        one = torch.zeros((K * rows, pci.cin_pad), device=dev, dtype=dtype)
        ops.stem_gather(frames[c:c + 1], masks[c], one)
        assert torch.equal(gi[c * K * rows:(c + 1) * K * rows], one)
    _lib.poll_kernel_error()


def pc3_weight_as_conv(pc, cin):
    """Invert the s2d packing back to a [cout, cin, 7, 7] kernel (float64): proves the packing is a permutation."""
    w = torch.zeros((pc.cout, cin, 7, 7), dtype=torch.float64, device=pc.weight.device)
    W = pc.weight.double()
    for t in range(4):
        for py in range(2):
            ky = 2 * (t - 2) + py + 3
            for dxi in range(4):
                for px in range(2):
                    kx = 2 * (dxi - 2) + px + 3
                    k0 = py * 8 * cin + dxi * 2 * cin + px * cin
                    if ky < 0 or kx < 0:
                        assert float(W[t, :, k0:k0 + cin].abs().max()) == 0.0
                    else:
                        w[:, :, ky, kx] = W[t, :pc.cout, k0:k0 + cin]
    return w


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_batched_forms_equal_per_clip_launches(dev, dtype):
    g = torch.Generator().manual_seed(5)
    C, K, h, w, c = 3, 2, 12, 20, 64
    N = C * K
    up = _halo(torch.randn((N, c, h // 2, w // 2), generator=g).to(dev), dtype)
    skip = _halo(torch.randn((C, c, h, w), generator=g).to(dev), dtype)
    x1, r1 = torch.zeros((N, h + 2, w + 2, c), device=dev, dtype=dtype), torch.zeros((N, h + 2, w + 2, c), device=dev, dtype=dtype)
    ops.upsample2x_add(x1, up, N, h, w, x_relu=r1, skip=skip)
    x2, r2 = torch.zeros_like(x1), torch.zeros_like(r1)
    for ci in range(C):
        o = slice(ci * K, (ci + 1) * K)
        ops.upsample2x_add(x2[o], up[o], K, h, w, x_relu=r2[o], skip=skip[ci:ci + 1])
    assert torch.equal(x1, x2) and torch.equal(r1, r2)
    # halo_copy: C source maps broadcast over K images each
    src = _halo(torch.randn((C, 96, h, w), generator=g).to(dev))
    d1 = torch.zeros((N, h + 2, w + 2, 128), device=dev, dtype=dtype)
    ops.halo_copy(src, d1, N, h, w, 64, src_coff=32, dst_coff=64, relu=True)
    for i in range(N):
        assert torch.equal(d1[i, 1:-1, 1:-1, 64:128].float(), src[i // K, 1:-1, 1:-1, 32:96].clamp_min(0).to(dtype).float())
    # aggregation groups
    lg = _halo(torch.randn((N, 32, h, w), generator=g).to(dev) * 3)
    p1 = torch.zeros((C, K + 1, 1, 4 * h, 4 * w), device=dev)
    ops.upsample4x_sigmoid_aggregate(lg, K, h, w, prob_out=p1, groups=C)
    for ci in range(C):
        _, p = ops.upsample4x_sigmoid_aggregate(lg[ci * K:(ci + 1) * K], K, h, w)
        assert torch.equal(p1[ci], p)
    _lib.poll_kernel_error()


@pytest.mark.parametrize("algo", [ops.MEMREAD_TCGEN05, ops.MEMREAD_EXACT_SIMT])
@pytest.mark.parametrize("C,K,T,hw,k", [(4, 1, 6, 30 * 54, 20), (2, 3, 3, 12 * 20, 50)])
def test_memory_read_query_sets_equal_per_clip_reads(dev, algo, C, K, T, hw, k):
    g = torch.Generator().manual_seed(C * 100 + K)
    N, slots = C * K, T * hw
    bk = torch.randn((N, slots + hw, 128), generator=g).to(dev)
    bv = torch.randn((N, slots + hw, 512), generator=g).to(dev)
    qk = torch.randn((C, hw, 128), generator=g).to(dev)
    out1 = torch.zeros((N, hw, 512), device=dev)
    ws = torch.empty(ops.memory_read_workspace_bytes(N, slots, hw, k), dtype=torch.uint8, device=dev)
    _, i1, v1 = ops.memory_read(bk, bv, slots, qk, k, out1, workspace=ws, algo=algo, want_topk=True, q_div=K)
    for ci in range(C):
        o = slice(ci * K, (ci + 1) * K)
        out2 = torch.zeros((K, hw, 512), device=dev)
        ws2 = torch.empty(ops.memory_read_workspace_bytes(K, slots, hw, k), dtype=torch.uint8, device=dev)
        _, i2, v2 = ops.memory_read(bk[o].contiguous(), bv[o].contiguous(), slots, qk[ci], k, out2, workspace=ws2, algo=algo, want_topk=True)
        assert torch.equal(i1[o], i2) and torch.equal(v1[o], v2) and torch.equal(out1[o], out2)
    _lib.poll_kernel_error()


@pytest.mark.parametrize("act", ["fp16", "tf32"])
def test_forward_and_backward_pass_as_two_lanes(dev, prop_sd, act):
    """A mid-clip interaction: the forward and the backward pass run as two concurrent lanes (own graphs, bank,
    workspace, stream) and must leave exactly what the sequential passes leave (reference order: inference_core.py:
    255-256), and agree with the oracle."""
    import mivos_b200
    from oracle import stm_oracle as O, weights as Wt
    net = mivos_b200.PropagationNetwork(top_k=20, act_dtype=torch.float16 if act == "fp16" else torch.float32)
    net.load_state_dict(prop_sd, strict=True)
    net = net.to(dev)
    images, mask = Wt.synthetic_clip(12, 96, 128, 2, seed=21)
    res = {}
    for overlap in (True, False):
        core = mivos_b200.InferenceCore(net, None, images, 2, mem_freq=2, device=dev)
        core.overlap_passes = overlap
        assert core._passes_can_overlap(5) == overlap
        steps = []
        m = core.interact(mask.to(dev), 5, total_cb=lambda n: steps.append(("total", n)), step_cb=lambda: steps.append("s"))
        torch.cuda.synchronize()
        _lib.poll_kernel_error()
        res[overlap] = (m, core.prob.clone(), list(core.bank_trace), steps)
    assert res[True][2] == res[False][2] and res[True][3] == res[False][3]
    assert torch.equal(res[True][1], res[False][1]) and (res[True][0] == res[False][0]).all()
    oc = O.OracleInferenceCore(prop_sd, None, images, 2, mem_freq=2, top_k=20)
    om = oc.interact(mask, 5)
    assert oc.bank_trace == res[True][2]
    assert float((res[True][1].cpu() - oc.prob).abs().max()) <= 3e-2 and float((res[True][0] != om).mean()) <= 0.01
