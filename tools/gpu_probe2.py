"""Network-level parity probe (run under gpurun): mivos_b200 vs the CPU oracle on the committed
golden fixtures, plus per-stage timings of one propagated 480p frame.  Prints, does not assert."""
import sys, os, time, json, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mivos_b200
from mivos_b200 import ops, _lib
from oracle import stm_oracle as O, weights as Wt

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
RES = {}


def rel(a, b):
    a, b = a.double().cpu(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max()), float(((a - b).pow(2).mean().sqrt()) / b.pow(2).mean().sqrt())


def run(fn, *a):
    try:
        fn(*a)
    except Exception as e:
        print("FAILED", fn.__name__, repr(e), flush=True)
        traceback.print_exc()


psd, fsd = Wt.make_prop_state_dict(), Wt.make_fusion_state_dict()
net = mivos_b200.PropagationNetwork(top_k=20)
net.load_state_dict(psd)
net = net.to(dev)
fuse = mivos_b200.FusionNet()
fuse.load_state_dict(fsd)
fuse = fuse.to(dev)


def ops_lowres():
    g = np.load(os.path.join(G, "ops_lowres.npz"))
    frame = torch.from_numpy(g["frame"]).to(dev)
    mask = torch.from_numpy(g["mask"]).to(dev)
    f16, f8, f4, k16, v16 = net.get_query_values(frame)
    print("qv f16", rel(f16, g["f16"]), "f8", rel(f8[:, ::4], g["f8"]), "f4", rel(f4[:, ::8], g["f4"]),
          "k16", rel(k16, g["k16"]), "v16", rel(v16, g["v16"]), flush=True)
    mk, mv = net.memorize(frame, mask[1:])
    print("memorize k", rel(mk, g["mem_k"]), "v", rel(mv, g["mem_v"]), flush=True)
    keys, values = torch.from_numpy(g["keys"]).to(dev), torch.from_numpy(g["values"]).to(dev)
    qv3 = net.get_query_values(torch.from_numpy(g["frame3"]).to(dev))
    print("qk3", rel(qv3[3], g["qk3"]), flush=True)
    # segment with ORACLE query features (isolates decoder+read error) and with our own
    ref_q = O.get_query_values(psd, torch.from_numpy(g["frame3"]))
    seg_a = net.segment_with_query(keys, values, *[t.to(dev) for t in ref_q])
    seg_b = net.segment_with_query(keys, values, *qv3)
    sg = torch.from_numpy(g["seg"])
    print("segment (oracle feats) max|dp|", float((seg_a.cpu() - sg).abs().max()), "mean", float((seg_a.cpu() - sg).abs().mean()),
          "| own feats max|dp|", float((seg_b.cpu() - sg).abs().max()), "mean", float((seg_b.cpu() - sg).abs().mean()), flush=True)
    ag = mivos_b200.aggregate_wbg(torch.from_numpy(g["seg"]).to(dev), keep_bg=True)
    print("aggregate max|d|", float((ag.cpu() - torch.from_numpy(g["agg"])).abs().max()), flush=True)
    at = net.get_attention(torch.from_numpy(g["mem_k"][0:1]).to(dev), torch.from_numpy(g["pos"]).to(dev),
                           torch.from_numpy(g["neg"]).to(dev), torch.from_numpy(g["qk3"]).to(dev))
    print("attention", rel(at, g["attn"]), flush=True)
    fu = fuse(torch.from_numpy(g["frame3"]).to(dev), torch.from_numpy(g["seg"][0:1]).to(dev),
              torch.from_numpy(g["agg"][1:2]).to(dev), torch.from_numpy(g["attn"]).to(dev), torch.from_numpy(g["dist"]).to(dev))
    print("fusion logit", rel(fu, g["fuse"]), flush=True)


def memread_golden():
    g = np.load(os.path.join(G, "memread.npz"))
    mk, mv, qk = (torch.from_numpy(g[n]).to(dev) for n in ("mk", "mv", "qk"))
    K, _, T, h, w = mk.shape
    hw = h * w
    bk = torch.empty((K, T * hw, 128), device=dev)
    bv = torch.empty((K, T * hw, 512), device=dev)
    ops.bank_from_nchw(mk, mv, bk, bv)
    qpm = qk.reshape(128, hw).t().contiguous()
    for k_ in (20, 50):
        out = torch.zeros((K, hw, 512), device=dev)
        ops.memory_read(bk, bv, T * hw, qpm, k_, out, algo=ops.MEMREAD_EXACT_SIMT)
        ref = torch.from_numpy(g[f"out{k_}"]).reshape(K, 512, hw).transpose(1, 2)
        print(f"memread golden top{k_}", rel(out, ref), flush=True)


def clip_lowres():
    g = np.load(os.path.join(G, "clip_lowres.npz"))
    images = torch.from_numpy(g["images"])
    core = mivos_b200.InferenceCore(net, fuse, images, 2, mem_profile=0, mem_freq=2, device="cuda:0")
    m1 = core.interact(torch.from_numpy(g["mask"]), 0)
    p1 = core.prob.cpu()
    print("clip interact1: mask mismatch frac", float((m1 != g["masks1"]).mean()), "prob max|d|", float((p1 - torch.from_numpy(g["prob1"])).abs().max()),
          "mean|d|", float((p1 - torch.from_numpy(g["prob1"])).abs().mean()), "trace", core.bank_trace, flush=True)
    core.bank_trace = []
    m2 = core.interact(torch.from_numpy(g["mask2"]), images.shape[1] - 1)
    p2 = core.prob.cpu()
    print("clip interact2 (fusion): mask mismatch frac", float((m2 != g["masks2"]).mean()), "prob max|d|",
          float((p2 - torch.from_numpy(g["prob2"])).abs().max()), "mean|d|", float((p2 - torch.from_numpy(g["prob2"])).abs().mean()),
          "trace", core.bank_trace, flush=True)


def cfg1_480p():
    g = np.load(os.path.join(G, "cfg1_480p.npz"))
    images, mask = Wt.synthetic_clip(5, 480, 854, 1, seed=1234)
    n50 = mivos_b200.PropagationNetwork(top_k=50)
    n50.load_state_dict(psd)
    n50 = n50.to(dev)
    core = mivos_b200.InferenceCore(n50, None, images, 1, mem_profile=0, mem_freq=2, device="cuda:0")
    t0 = time.time()
    m = core.interact(mask, 0)
    torch.cuda.synchronize()
    t1 = time.time()
    print("cfg1 480p: mask mismatch frac", float((m != g["masks"]).mean()), "fg frac", float((m > 0).mean()), "ref fg", float((g["masks"] > 0).mean()),
          "prob_sub max|d|", float((core.prob[:, :, :, ::8, ::8].cpu() - torch.from_numpy(g["prob_sub"])).abs().max()),
          "first-call seconds", t1 - t0, "trace", core.bank_trace, flush=True)
    # per-stage timing of a steady-state frame
    K, hw = 1, 30 * 54
    T = 20
    n20 = net
    frame = core.images[:, 1]
    qs = n20.encode_query_resident(frame)
    bank_k = torch.randn((K, T * hw + hw, 128), device=dev)
    bank_v = torch.randn((K, T * hw + hw, 512), device=dev)
    mask1 = torch.rand((1, 1, 480, 864), device=dev)

    def timeit(name, fn, iters=10):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        c0 = _lib.load().mivos_launch_count()
        t0 = time.time()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.time() - t0) / iters * 1e3
        print(f"  stage {name}: gpu {e0.elapsed_time(e1)/iters:.3f} ms, wall {wall:.3f} ms, launches {(_lib.load().mivos_launch_count()-c0)//iters}", flush=True)

    timeit("encode_query", lambda: n20.encode_query_resident(frame, qs))
    timeit("segment(T=20,exact-simt read)", lambda: n20.segment_resident(bank_k, bank_v, T * hw, qs, K))
    timeit("memorize", lambda: n20.memorize_resident(frame, mask1, bank_k, bank_v, T))
    ws = torch.empty(ops.memory_read_workspace_bytes(K, T * hw, hw, 20), dtype=torch.uint8, device=dev)
    out = torch.zeros((K, hw, 512), device=dev)
    timeit("memory_read exact only", lambda: ops.memory_read(bank_k, bank_v, T * hw, qs.qk, 20, out, workspace=ws, algo=ops.MEMREAD_EXACT_SIMT))


run(ops_lowres)
run(memread_golden)
run(clip_lowres)
run(cfg1_480p)
_lib.poll_kernel_error()
print("done", flush=True)
