"""tcgen05 memory-read probe: bit-equality with the exact CUDA-core path, candidate statistics,
overflow fallback, ragged shapes, timings (run under gpurun)."""
import sys, os, time, traceback, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mivos_b200 import ops, _lib

dev = torch.device("cuda:0")
torch.manual_seed(0)


def case(name, K, T, h, w, top_k, mutate=None, time_it=False, extra_slots=0):
    hw = h * w
    slots = T * hw + extra_slots
    cap = slots + 300
    bk = torch.randn((K, cap, 128), device=dev)
    bv = torch.randn((K, cap, 512), device=dev)
    qk = torch.randn((hw, 128), device=dev)
    if mutate:
        mutate(bk, bv, qk, slots)
    ws = torch.empty(ops.memory_read_workspace_bytes(K, slots, hw, top_k), dtype=torch.uint8, device=dev)
    o1 = torch.zeros((K, hw, 512), device=dev)
    o2 = torch.zeros((K, hw, 512), device=dev)
    _, i1, v1 = ops.memory_read(bk, bv, slots, qk, top_k, o1, workspace=ws, algo=ops.MEMREAD_EXACT_SIMT, want_topk=True)
    torch.cuda.synchronize()
    try:
        _, i2, v2 = ops.memory_read(bk, bv, slots, qk, top_k, o2, workspace=ws, algo=ops.MEMREAD_TCGEN05, want_topk=True)
        torch.cuda.synchronize()
        _lib.poll_kernel_error()
    except Exception as e:
        print(name, "TC FAILED", repr(e), flush=True)
        return
    st = (C.c_int64 * 4)()
    _lib.check(_lib.lib().mivos_memory_read_stats(C.c_void_p(ws.data_ptr()), K, slots, hw, top_k, st))
    if K * slots * hw <= 40_000_000:
        # torch fp64 reference (the reference's own formulation) to tell WHICH path is off
        aff = torch.einsum("ksc,qc->ksq", bk[:, :slots].double(), (qk / (128 ** 0.5)).double())
        vals, ind = torch.topk(aff, top_k, dim=1)
        for nm, ii in (("exact", i1), ("tc", i2)):
            ok = (ii.long().transpose(1, 2).sort(1)[0] == ind.sort(1)[0]).all(1)
            bad = (~ok).nonzero()
            print(f"   vs torch: {nm} index sets equal {int(ok.sum())}/{ok.numel()}; first bad (obj,q): {bad[:6].tolist()}", flush=True)
    same_idx = (i1 == i2).all(-1)
    print(f"{name}: K={K} slots={slots} hw={hw} k={top_k} | idx equal {int(same_idx.sum())}/{same_idx.numel()} | scores equal {bool((v1 == v2).all())} | "
          f"readout bit-equal {bool(torch.equal(o1, o2))} max|d| {(o1 - o2).abs().max().item():.2e} | cand/query mean {st[0] / (K * hw - st[2] + 1e-9):.1f} max {st[1]} "
          f"flagged {st[2]} splits {st[3]}", flush=True)
    if time_it:
        for algo, nm in ((ops.MEMREAD_TCGEN05, "tcgen05"), (ops.MEMREAD_EXACT_SIMT, "exact")):
            for _ in range(3):
                ops.memory_read(bk, bv, slots, qk, top_k, o2, workspace=ws, algo=algo)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            n = 10
            for _ in range(n):
                ops.memory_read(bk, bv, slots, qk, top_k, o2, workspace=ws, algo=algo)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / n * 1e3
            print(f"   {nm}: {us:.1f} us  ({2*128*slots*hw*K/us/1e6:.1f} TFLOP/s on QK^T)", flush=True)


def run(*a, **k):
    try:
        case(*a, **k)
    except Exception as e:
        print("FAILED", a, repr(e), flush=True)
        traceback.print_exc()


def all_equal_keys(bk, bv, qk, slots):
    bk[0, :] = bk[0, 0]


def duplicates(bk, bv, qk, slots):
    bk[:, 1:slots:2] = bk[:, 0:slots - 1:2]  # every key appears twice: exact ties


def big_norm(bk, bv, qk, slots):
    bk[:, 5] *= 50.0  # one huge key inflates the margin


run("T1", 1, 1, 30, 54, 20)
run("T1_small", 1, 1, 6, 8, 20)
run("T2_hw40", 1, 2, 5, 8, 20)
run("T3", 1, 3, 30, 54, 20)
run("T5_k50_K2", 2, 5, 30, 54, 50)
run("ragged_hw63", 2, 9, 7, 9, 20, extra_slots=5)
run("ragged_hw200_k50", 1, 13, 10, 20, 50, extra_slots=77)
run("dup_ties", 1, 4, 30, 54, 20, mutate=duplicates)
run("all_equal", 2, 3, 30, 54, 20, mutate=all_equal_keys)
run("big_norm", 1, 4, 30, 54, 20, mutate=big_norm)
run("cfg2_T20_k20", 1, 20, 30, 54, 20, time_it=True)
print("done", flush=True)
