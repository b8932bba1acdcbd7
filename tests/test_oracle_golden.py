"""CPU: the oracle restatement replayed against the fixtures generated from the UNMODIFIED
reference (oracle/gen_golden.py, tests/golden/MANIFEST.json).  Same torch build => bit-exact; the
tolerance 1e-5 only absorbs a different CPU's oneDNN/MKL kernel choice on the GPU box host."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import stm_oracle as O, weights as Wt

TOL = 1e-5


def _close(a, b, tol=TOL):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    assert a.shape == b.shape
    assert float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max())), float((a - b).abs().max())


def test_manifest_says_oracle_equals_reference(golden):
    m = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "MANIFEST.json")))
    assert m["prop_keys"] == 597 and m["fusion_keys"] == 12
    diffs = [v["max_abs_diff_oracle_vs_reference"] for v in m.values() if isinstance(v, dict) and "ref_abs_max" in v]
    assert len(diffs) >= 15 and max(diffs) == 0.0  # bit-identical to the reference when generated
    assert m["clip_masks_equal"] == [True, True] and m["cfg1_masks_equal"] is True


def test_state_dict_tables_match_reference_counts(prop_sd, fuse_sd):
    assert len(prop_sd) == 597 and len(fuse_sd) == 12
    assert prop_sd["mask_rgb_encoder.conv1.weight"].shape == (64, 5, 7, 7)
    assert "rgb_encoder.conv1.bias" not in prop_sd and "mask_rgb_encoder.conv1.bias" in prop_sd


def test_query_and_memorize(golden, prop_sd):
    g = golden("ops_lowres.npz")
    frame, mask = torch.from_numpy(g["frame"]), torch.from_numpy(g["mask"])
    f16, f8, f4, k16, v16 = O.get_query_values(prop_sd, frame)
    _close(f16, g["f16"]); _close(f8[:, ::4], g["f8"]); _close(f4[:, ::8], g["f4"])
    _close(k16, g["k16"]); _close(v16, g["v16"])
    mk, mv = O.memorize(prop_sd, frame, mask[1:])
    _close(mk, g["mem_k"]); _close(mv, g["mem_v"])


def test_segment_aggregate_attention_fusion(golden, prop_sd, fuse_sd):
    g = golden("ops_lowres.npz")
    keys, values = torch.from_numpy(g["keys"]), torch.from_numpy(g["values"])
    qv = O.get_query_values(prop_sd, torch.from_numpy(g["frame3"]))
    _close(O.memory_read(keys[0:1], values[0:1], qv[3], 20), g["read"])
    seg = O.segment_with_query(prop_sd, keys, values, *qv, top_k=20)
    _close(seg, g["seg"])
    _close(O.aggregate_wbg(seg, keep_bg=True), g["agg"])
    at = O.get_attention(None, torch.from_numpy(g["mem_k"][0:1]), torch.from_numpy(g["pos"]), torch.from_numpy(g["neg"]), qv[3])
    _close(at, g["attn"])
    fu = O.fusion_net(fuse_sd, torch.from_numpy(g["frame3"]), torch.from_numpy(g["seg"][0:1]), torch.from_numpy(g["agg"][1:2]),
                      torch.from_numpy(g["attn"]), torch.from_numpy(g["dist"]))
    _close(fu, g["fuse"], 1e-4)


@pytest.mark.parametrize("k", [20, 50])
def test_memory_read_fp32_and_f64(golden, k):
    g = golden("memread.npz")
    mk, mv, qk = (torch.from_numpy(g[n]) for n in ("mk", "mv", "qk"))
    out = torch.cat([O.memory_read(mk[i:i + 1], mv[i:i + 1], qk, k) for i in range(2)], 0)
    _close(out, g[f"out{k}"])
    # independent float64 statement agrees with the reference's fp32 result to fp32 rounding
    o64, idx, gap = O.memory_read_f64(g["mk"], g["mv"], g["qk"], k)
    _close(o64.reshape(out.shape), g[f"out{k}"], 2e-5)
    assert idx.shape == (2, k, 48) and (gap > 0).all()


def test_clip_lowres_with_fusion(golden, prop_sd, fuse_sd):
    g = golden("clip_lowres.npz")
    images = torch.from_numpy(g["images"])
    core = O.OracleInferenceCore(prop_sd, fuse_sd, images, 2, mem_freq=2, top_k=20)
    m1 = core.interact(torch.from_numpy(g["mask"]), 0)
    assert (m1 == g["masks1"]).all()
    _close(core.prob, g["prob1"])
    # bank bookkeeping of inference_core.py:165-186 with mem_freq=2 on 6 frames
    assert core.bank_trace == [(1, 1), (2, 2), (3, 2), (4, 3), (5, 3)]
    core.bank_trace = []
    m2 = core.interact(torch.from_numpy(g["mask2"]), images.shape[1] - 1)
    assert (m2 == g["masks2"]).all()
    _close(core.prob, g["prob2"])
    assert core.bank_trace == [(4, 2), (3, 3), (2, 3), (1, 4)]
    assert tuple(g["pad"]) == tuple(core.pad) == (4, 4, 0, 0)


def test_pad_and_aggregate_edge_cases():
    x = torch.arange(2 * 3 * 5 * 7, dtype=torch.float32).reshape(2, 3, 5, 7)
    y, pad = O.pad_divide_by(x, 16)
    assert y.shape[-2:] == (16, 16) and pad == (4, 5, 5, 6)
    assert torch.equal(O.unpad(y, pad), x)
    y2, pad2 = O.pad_divide_by(torch.zeros(1, 1, 32, 48), 16)
    assert pad2 == (0, 0, 0, 0) and y2.shape[-2:] == (32, 48)
    p = torch.tensor([0.0, 1.0, 0.5, 1e-9]).view(1, 1, 2, 2)
    a = O.aggregate_wbg(p, keep_bg=True)
    assert torch.isfinite(a).all() and torch.allclose(a.sum(0), torch.ones(1, 2, 2))
    hard = O.aggregate_wbg(torch.tensor([0.6, 0.4]).view(2, 1, 1, 1), keep_bg=True, hard=True)
    assert int(hard.argmax(0)) == 1


def test_oracle_attention_read_network_matches_reference_golden(golden, prop_sd):
    """SURVEY §8(f) row 2 — oracle restatement of AttentionReadNetwork.forward vs the reference run."""
    import torch
    from oracle import stm_oracle as O
    g = golden("attn_read.npz")
    t = lambda n: torch.from_numpy(g[n])  # noqa: E731
    with torch.no_grad():
        a1, a2 = O.attention_read_network(prop_sd, t("image"), t("m11"), t("m21"), t("m12"), t("m22"), t("query"))
    assert float((a1 - t("attn1")).abs().max()) <= 1e-6 and float((a2 - t("attn2")).abs().max()) <= 1e-6


def test_float64_restatements_of_aggregate_and_attention(golden):
    """independent float64 numpy statements of a6 (aggregate_wbg) and a10 (get_attention)
    agree with the reference's fp32 outputs to fp32 rounding — they arbitrate when a GPU result and the fp32
    oracle disagree in the last bits."""
    g = golden("ops_lowres.npz")
    a = O.aggregate_wbg_f64(g["seg"], keep_bg=True)
    assert a.shape == g["agg"].shape and float(np.abs(a - g["agg"]).max()) <= 5e-7
    assert float(np.abs(a.sum(0) - 1).max()) <= 1e-12
    hard = O.aggregate_wbg_f64(g["seg"], keep_bg=True, hard=True)
    ref_hard = O.aggregate_wbg(torch.from_numpy(g["seg"]), keep_bg=True, hard=True).numpy()
    decided = np.abs(hard.max(0) - 1) < 1e-9  # away from exact ties the x1000 softmax is a one-hot
    assert decided.mean() > 0.95 and (hard.argmax(0) == ref_hard.argmax(0))[decided].all()
    at = O.get_attention_f64(g["mem_k"][0:1], g["pos"], g["neg"], g["qk3"])
    assert at.shape == g["attn"].shape
    assert float(np.abs(at - g["attn"]).max()) <= 2e-6 * max(1.0, float(np.abs(g["attn"]).max()))
