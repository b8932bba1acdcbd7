"""CPU: property tests (hypothesis) of the host arithmetic the runtime relies on — size-independent
invariants rather than fixtures: the pass plan against a literal simulation of the reference's loop
(inference_core.py:122-200), pad / unpad round trips (util/tensor_util.py:62-87), the indexed-PNG
writer against an independent decoder."""
import io

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from mivos_b200 import egress, schedule, tensor_util


def _reference_loop(t, interacted, idx, forward, mem_freq, nck):
    """inference_core.py:128-186 with the tensors removed: returns (closest_ti, total_m, trace) where
    trace = [(ti, slots the read sees, slot memorize writes or None)]."""
    m_front = nck
    if forward:
        closest_ti = min([ti for ti in interacted if ti > idx] + [t])
        total_m = (closest_ti - idx - 1) // mem_freq + 1 + nck
        this_range, end = range(idx + 1, closest_ti), closest_ti - 1
    else:
        closest_ti = max([ti for ti in interacted if ti < idx] + [-1])
        total_m = (idx - closest_ti - 1) // mem_freq + 1 + nck
        this_range, end = range(idx - 1, closest_ti, -1), closest_ti + 1
    prev_in_mem, last_ti, trace = True, idx, []
    for ti in this_range:
        seen = m_front if prev_in_mem else m_front + 1
        wrote = None
        if ti != end:
            wrote = m_front
            if abs(ti - last_ti) >= mem_freq:
                m_front += 1
                last_ti = ti
                prev_in_mem = True
            else:
                prev_in_mem = False
        trace.append((ti, seen, wrote))
    return closest_ti, total_m, trace


@settings(max_examples=300, deadline=None)
@given(t=st.integers(1, 60), mem_freq=st.integers(1, 9), data=st.data())
def test_plan_pass_equals_the_reference_loop(t, mem_freq, data):
    interacted = set(data.draw(st.lists(st.integers(0, t - 1), min_size=1, max_size=5)))
    idx = data.draw(st.sampled_from(sorted(interacted)))
    nck = len(interacted)
    for forward in (True, False):
        closest, total_m, trace = _reference_loop(t, interacted, idx, forward, mem_freq, nck)
        plan = schedule.plan_pass(t, interacted, idx, forward, mem_freq, nck)
        assert (plan.closest_ti, plan.total_m) == (closest, total_m)
        assert [(f.ti, f.visible, f.m_front if f.memorize else None) for f in plan.frames] == trace
        assert plan.fuse == (closest not in (-1, t))
        # the bank the pass allocates is never overrun, and reads never see an unwritten slot
        written = set(range(nck))
        for f in plan.frames:
            assert f.visible <= plan.total_m and set(range(f.visible)) <= written
            if f.memorize:
                assert f.m_front < plan.total_m
                written.add(f.m_front)
        assert schedule.bank_capacity_frames(t, mem_freq, nck, plan.total_m) >= plan.total_m


@settings(max_examples=100, deadline=None)
@given(h=st.integers(1, 70), w=st.integers(1, 70), d=st.sampled_from([2, 4, 16]))
def test_pad_unpad_round_trip(h, w, d):
    x = torch.arange(2 * 3 * h * w, dtype=torch.float32).reshape(2, 3, h, w)
    y, pad = tensor_util.pad_divide_by(x, d)
    assert y.shape[-2] % d == 0 and y.shape[-1] % d == 0 and y.shape[-2] - h < d and y.shape[-1] - w < d
    assert pad == tensor_util.pad_amounts(h, w, d) and pad[0] + pad[1] == y.shape[-1] - w and pad[2] + pad[3] == y.shape[-2] - h
    assert abs(pad[0] - pad[1]) <= 1 and abs(pad[2] - pad[3]) <= 1 and pad[0] <= pad[1] and pad[2] <= pad[3]  # symmetric, extra on the far side
    assert torch.equal(tensor_util.unpad(y, pad), x)
    assert float(y.double().sum()) == float(x.double().sum())  # zero padding


@settings(max_examples=40, deadline=None)
@given(h=st.integers(1, 64), w=st.integers(1, 64), labels=st.integers(1, 255), seed=st.integers(0, 2**31 - 1))
def test_indexed_png_decodes_to_the_same_labels(h, w, labels, seed):
    Image = pytest.importorskip("PIL.Image")
    mask = np.random.default_rng(seed).integers(0, labels + 1, size=(h, w), dtype=np.uint8)
    png = egress.encode_indexed_png(mask)
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    im = Image.open(io.BytesIO(png))
    im.load()
    assert im.mode == "P" and np.array_equal(np.asarray(im), mask)
