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
