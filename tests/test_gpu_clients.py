"""GPU: the second client of the drop-in surface — the reference's
FusionGenerator (generation/fusion_generator.py:40-100) drives PropagationNetwork through its
REFERENCE-LAYOUT methods only (memorize / get_query_values / segment_with_query with banks grown
by torch.cat, aggregate_wbg, pad_divide_by).  The flow below is that client restated as test code
and run twice: on mivos_b200 (CUDA) and on the CPU oracle; the soft probabilities must agree
within the TF32 tolerance of test_gpu_network.py."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

import mivos_b200  # noqa: E402
from mivos_b200 import _lib  # noqa: E402
from oracle import stm_oracle as O, weights as Wt  # noqa: E402  (checker only)


def fusion_generator_flow(api, images, mask, idx, left_limit, right_limit, mem_freq):
    """generation/fusion_generator.py:82-100 (interact_mask) + :40-80 (do_pass)."""
    images, pad = api.pad_divide_by(images, 16, images.shape[-2:])
    t = images.shape[1]
    nh, nw = images.shape[-2:]
    mask, _ = api.pad_divide_by(mask, 16, mask.shape[-2:])
    k = mask.shape[0]
    prob = torch.zeros((k + 1, t, 1, nh, nw), dtype=torch.float32, device=images.device)
    m = api.aggregate_wbg(mask, keep_bg=True)
    prob[:, idx] = m
    key_k, key_v = api.memorize(images[:, idx], m[1:])
    for forward in (True, False):
        keys, values, prev_k, prev_v, last_ti = key_k, key_v, None, None, idx
        rng, end = (range(idx + 1, right_limit + 1), right_limit) if forward else (range(idx - 1, left_limit - 1, -1), left_limit)
        for ti in rng:
            this_k = keys if prev_k is None else torch.cat([keys, prev_k], 2)
            this_v = values if prev_v is None else torch.cat([values, prev_v], 2)
            query = api.get_query_values(images[:, ti])
            out = api.aggregate_wbg(api.segment_with_query(this_k, this_v, *query), keep_bg=True)
            prob[:, ti] = out
            if ti != end:
                prev_k, prev_v = api.memorize(images[:, ti], out[1:])
                if abs(ti - last_ti) >= mem_freq:
                    last_ti = ti
                    keys, values = torch.cat([keys, prev_k], 2), torch.cat([values, prev_v], 2)
                    prev_k = prev_v = None
    return prob


def test_fusion_generator_client(dev, nets, prop_sd):
    net = nets[50]  # generate_fusion.py builds PropagationNetwork(top_k=50)
    # 168 -> padded to 176; 8 x 11 = 88 bank slots per frame >= top_k (torch.topk in the reference
    # raises as well when the bank holds fewer than k slots)
    images, mask = Wt.synthetic_clip(6, 128, 168, 2, seed=31)
    soft = mask[1:] * 0.8 + 0.05
    ours = types.SimpleNamespace(pad_divide_by=mivos_b200.pad_divide_by, aggregate_wbg=mivos_b200.aggregate_wbg,
                                 memorize=net.memorize, get_query_values=net.get_query_values,
                                 segment_with_query=net.segment_with_query)
    ref = types.SimpleNamespace(pad_divide_by=O.pad_divide_by, aggregate_wbg=O.aggregate_wbg,
                                memorize=lambda f, m: O.memorize(prop_sd, f, m),
                                get_query_values=lambda f: O.get_query_values(prop_sd, f),
                                segment_with_query=lambda *a: O.segment_with_query(prop_sd, *a, top_k=50))
    p_gpu = fusion_generator_flow(ours, images.to(dev), soft.to(dev), 2, 0, 5, mem_freq=2).cpu()
    p_cpu = fusion_generator_flow(ref, images, soft, 2, 0, 5, mem_freq=2)
    d = (p_gpu - p_cpu).abs()
    assert float(d.max()) <= 3e-2 and float(d.mean()) <= 1e-3
    _lib.poll_kernel_error()


def test_attention_read_network_matches_reference_golden(dev, golden, prop_sd):
    """AttentionReadNetwork.forward (model/attn_network.py:48-80), loaded from a
    propagation state dict with strict=False like model/fusion_model.py:187; golden maps come from
    the unmodified reference module (oracle/gen_golden_attn.py, oracle == reference: 0.0)."""
    g = golden("attn_read.npz")
    for act in (torch.float16, torch.float32):
        net = mivos_b200.AttentionReadNetwork(act_dtype=act)
        missing, unexpected = net.load_state_dict(prop_sd, strict=False)
        assert not missing and all(k.startswith("decoder.") for k in unexpected)
        net = net.to(dev)
        t = lambda n: torch.from_numpy(g[n]).to(dev)  # noqa: E731
        a1, a2 = net(t("image"), t("m11"), t("m21"), t("m12"), t("m22"), t("query"))
        for got, name in ((a1, "attn1"), (a2, "attn2")):
            ref = torch.from_numpy(g[name])
            assert got.shape == ref.shape
            d = (got.cpu() - ref).abs()
            # softmax over 24 memory positions of conv-stack keys: the conv tolerance (4e-3 of the
            # feature range) bounds the logit error; maps are convex combinations of pooled masks
            assert float(d.max()) <= 2e-2 * float(ref.abs().max()) + 1e-4, float(d.max())
    _lib.poll_kernel_error()


def test_ingest_is_bit_identical(dev):
    """images_to_torch (interact/interactive_utils.py:18-23) — u8 HWC frames ->
    normalised fp32 NCHW; the reference's arithmetic restated on the CPU, compared bit for bit."""
    from mivos_b200.ingest import images_to_torch
    g = torch.Generator().manual_seed(3)
    frames = torch.randint(0, 256, (5, 37, 53, 3), generator=g, dtype=torch.uint8)
    frames[0, 0, 0] = torch.tensor([0, 255, 128], dtype=torch.uint8)
    got = images_to_torch(frames.numpy(), dev)
    ref = frames.permute(0, 3, 1, 2).float().unsqueeze(0) / 255          # :19
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 1, 3, 1, 1)        # dataset/range_transform.py:5-8
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 1, 3, 1, 1)
    ref = (ref - mean) / std                                              # torchvision Normalize: sub_ then div_
    assert got.shape == (1, 5, 3, 37, 53) and got.dtype == torch.float32
    assert torch.equal(got.cpu(), ref)
    _lib.poll_kernel_error()
