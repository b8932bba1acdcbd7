"""First-contact probe for the sm_100a kernels (run under gpurun). Prints diagnostics rather than
asserting, so one GPU call answers as many open questions as possible (TF32 rounding mode of
tcgen05, descriptor correctness for every tile width, elementwise kernels, exact memory read)."""
import sys, os, time, json, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mivos_b200 import ops, _lib

dev = torch.device("cuda:0")
torch.manual_seed(0)
RES = {}


def to_halo(x, cpad=None):
    n, c, h, w = x.shape
    cp = cpad or c
    hb = torch.zeros((n, h + 2, w + 2, cp), device=x.device, dtype=torch.float32)
    hb[:, 1:-1, 1:-1, :c] = x.permute(0, 2, 3, 1)
    return hb


def from_halo(hb, c):
    return hb[:, 1:-1, 1:-1, :c].permute(0, 3, 1, 2).contiguous()


def tf32_trunc(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def tf32_rna(x):
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def conv_case(name, n, h, w, cin, cout, ks, relu=False, residual=False, dual=False):
    x = torch.randn(n, cin, h, w, device=dev)
    wt = torch.randn(cout, cin, ks, ks, device=dev) / (cin * ks * ks) ** 0.5
    b = torch.randn(cout, device=dev)
    pc = ops.pack_conv(wt, b, device=dev)
    xin = to_halo(x, pc.cin_pad)
    out = torch.full((n, h + 2, w + 2, pc.cout_pad), 7.0, device=dev)  # sentinel: halo must stay 7
    res = torch.randn(n, cout, h, w, device=dev) if residual else None
    res_h = to_halo(res, pc.cout_pad) if residual else None
    out2 = torch.zeros_like(out) if dual else None
    ops.conv_gemm(xin, pc, n, h, w, out, relu=relu, residual=res_h, out_relu=out2)
    torch.cuda.synchronize()
    _lib.poll_kernel_error()
    got = from_halo(out, cout)
    refs = {}
    for nm, f in (("trunc", tf32_trunc), ("rna", tf32_rna), ("fp32", lambda t: t)):
        r = F.conv2d(f(x).double(), f(wt).double(), b.double(), padding=ks // 2)
        if residual:
            r = r + res.double()
        if relu:
            r = r.relu()
        refs[nm] = r
    scale = refs["fp32"].abs().max().item()
    errs = {nm: ((got.double() - r).abs().max().item() / scale) for nm, r in refs.items()}
    halo_ok = bool((out[:, 0] == 7).all() and (out[:, -1] == 7).all() and (out[:, :, 0] == 7).all() and (out[:, :, -1] == 7).all())
    pad_ok = bool((out[:, 1:-1, 1:-1, cout:] == 7).all()) if pc.cout_pad > cout else True
    d2 = None
    if dual:
        d2 = (from_halo(out2, cout) - got.relu()).abs().max().item()
    RES[name] = dict(err=errs, halo_ok=halo_ok, pad_ok=pad_ok, dual_err=d2)
    print(name, json.dumps(RES[name]), flush=True)


def run(fn, *a, **k):
    try:
        fn(*a, **k)
    except Exception as e:  # keep going: later probes are still informative
        print("FAILED", fn.__name__, a, repr(e), flush=True)
        traceback.print_exc()


print("device", torch.cuda.get_device_name(0), flush=True)
l = _lib.lib()
print("abi", l.mivos_abi_version(), flush=True)

run(conv_case, "c3x3_bn64", 1, 30, 54, 64, 64, 3)
run(conv_case, "c3x3_bn32_cout1", 1, 30, 54, 32, 1, 3)
run(conv_case, "c1x1_bn64_relu", 1, 30, 54, 256, 64, 1, relu=True)
run(conv_case, "c3x3_bn128_res", 2, 60, 108, 128, 128, 3, residual=True, relu=True)
run(conv_case, "c3x3_bn256_dual", 1, 120, 216, 256, 256, 3, dual=True)
run(conv_case, "c1x1_bn256_k1024", 1, 60, 108, 1024, 512, 1)
run(conv_case, "c3x3_cin1024_640", 1, 30, 54, 1024, 640, 3)


# ------------------------------------------------------------------ timing of big shapes
def time_conv(name, n, h, w, cin, cout, ks, iters=20):
    x = torch.randn(n, cin, h, w, device=dev)
    wt = torch.randn(cout, cin, ks, ks, device=dev) / (cin * ks * ks) ** 0.5
    pc = ops.pack_conv(wt, torch.zeros(cout, device=dev), device=dev)
    xin = to_halo(x, pc.cin_pad)
    out = torch.zeros((n, h + 2, w + 2, pc.cout_pad), device=dev)
    for _ in range(3):
        ops.conv_gemm(xin, pc, n, h, w, out)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        ops.conv_gemm(xin, pc, n, h, w, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * n * h * w * cin * ks * ks * cout
    print(f"time {name}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s (algorithmic)", flush=True)
    RES["time_" + name] = dict(us=ms * 1e3, tflops=fl / ms / 1e9)


run(time_conv, "dec_3x3_256_256_120x216", 1, 120, 216, 256, 256, 3)
run(time_conv, "dec_3x3_512_512_60x108", 1, 60, 108, 512, 512, 3)
run(time_conv, "l3_1x1_1024_256_30x54", 1, 30, 54, 1024, 256, 1)
run(time_conv, "l3_3x3_256_256_30x54", 1, 30, 54, 256, 256, 3)
run(time_conv, "fuse_3x3_32_32_480x864", 1, 480, 864, 32, 32, 3)


# ------------------------------------------------------------------ elementwise kernels
def elementwise():
    n, c, h, w = 2, 64, 60, 108
    x = torch.randn(n, c, h, w, device=dev).relu()
    xh = to_halo(x)
    out = torch.zeros((n, h // 2 + 2, w // 2 + 2, c), device=dev)
    ops.maxpool3x3s2(xh, n, h, w, out)
    ref = F.max_pool2d(x, 3, 2, 1)
    print("maxpool err", (from_halo(out, c) - ref).abs().max().item(), flush=True)
    # gather_s2 3x3 + conv == strided conv
    wt = torch.randn(96, c, 3, 3, device=dev) / 24
    b = torch.randn(96, device=dev)
    pc = ops.pack_conv(wt, b, stride=2, im2col=True, device=dev)
    g = torch.zeros((n * (h // 2 + 2) * (w // 2 + 2), pc.cin_pad), device=dev)
    ops.gather_s2(xh, n, h, w, c, 3, g)
    o = torch.zeros((n, h // 2 + 2, w // 2 + 2, pc.cout_pad), device=dev)
    ops.conv_gemm(g, pc, n, h // 2, w // 2, o)
    ref = F.conv2d(x.double(), wt.double(), b.double(), stride=2, padding=1)
    print("s2 3x3 conv rel err", ((from_halo(o, 96).double() - ref).abs().max() / ref.abs().max()).item(), flush=True)
    wt1 = torch.randn(128, c, 1, 1, device=dev) / 8
    pc1 = ops.pack_conv(wt1, None, stride=2, im2col=True, device=dev)
    g1 = torch.zeros((n * (h // 2 + 2) * (w // 2 + 2), pc1.cin_pad), device=dev)
    ops.gather_s2(xh, n, h, w, c, 1, g1)
    o1 = torch.zeros((n, h // 2 + 2, w // 2 + 2, pc1.cout_pad), device=dev)
    ops.conv_gemm(g1, pc1, n, h // 2, w // 2, o1)
    ref = F.conv2d(x.double(), wt1.double(), None, stride=2)
    print("s2 1x1 conv rel err", ((from_halo(o1, 128).double() - ref).abs().max() / ref.abs().max()).item(), flush=True)
    # stem
    H, W = 96, 160
    fr = torch.randn(1, 3, H, W, device=dev)
    mk = torch.rand(3, 1, H, W, device=dev)
    ws = torch.randn(64, 5, 7, 7, device=dev) / 15
    bs = torch.randn(64, device=dev)
    pcs = ops.pack_conv(ws, bs, stride=2, im2col=True, device=dev)
    gs = torch.zeros((3 * (H // 2 + 2) * (W // 2 + 2), pcs.cin_pad), device=dev)
    ops.stem_gather(fr, mk, gs)
    os_ = torch.zeros((3, H // 2 + 2, W // 2 + 2, 64), device=dev)
    ops.conv_gemm(gs, pcs, 3, H // 2, W // 2, os_, relu=True)
    others = torch.stack([mk.sum(0) - mk[i] for i in range(3)], 0)
    inp = torch.cat([fr.expand(3, -1, -1, -1), mk, others], 1)
    ref = F.conv2d(inp.double(), ws.double(), bs.double(), stride=2, padding=3).relu()
    print("stem5 rel err", ((from_halo(os_, 64).double() - ref).abs().max() / ref.abs().max()).item(), flush=True)
    pcs3 = ops.pack_conv(ws[:, :3].contiguous(), bs, stride=2, im2col=True, device=dev)
    gs3 = torch.zeros(((H // 2 + 2) * (W // 2 + 2), pcs3.cin_pad), device=dev)
    ops.stem_gather(fr, None, gs3)
    os3 = torch.zeros((1, H // 2 + 2, W // 2 + 2, 64), device=dev)
    ops.conv_gemm(gs3, pcs3, 1, H // 2, W // 2, os3)
    ref = F.conv2d(fr.double(), ws[:, :3].double(), bs.double(), stride=2, padding=3)
    print("stem3 rel err", ((from_halo(os3, 64).double() - ref).abs().max() / ref.abs().max()).item(), flush=True)
    # upsample2x_add
    a = torch.randn(n, c, h, w, device=dev)
    u = torch.randn(n, c, h // 2, w // 2, device=dev)
    ah, uh = to_halo(a), to_halo(u)
    ar = torch.zeros_like(ah)
    ops.upsample2x_add(ah, uh, n, h, w, x_relu=ar)
    ref = a + F.interpolate(u, scale_factor=2, mode="bilinear", align_corners=False)
    print("upsample2x_add err", (from_halo(ah, c) - ref).abs().max().item(), "relu err", (from_halo(ar, c) - ref.relu()).abs().max().item(), flush=True)
    # layout round trip
    back = ops.halo_to_nchw(xh, n, h, w, c)
    print("halo_to_nchw exact", bool((back == x).all()), flush=True)
    h2 = torch.zeros_like(xh)
    ops.nchw_to_halo(x, h2)
    print("nchw_to_halo exact", bool((h2 == xh).all()), flush=True)
    # upsample4x + sigmoid + aggregate
    k, h4, w4 = 3, 24, 40
    lg = torch.randn(k, 1, h4, w4, device=dev) * 3
    lh = to_halo(lg, 32)
    raw, prob = ops.upsample4x_sigmoid_aggregate(lh, k, h4, w4, want_raw=True)
    r = torch.sigmoid(F.interpolate(lg, scale_factor=4, mode="bilinear", align_corners=False))
    print("up4 sigmoid err", (raw - r).abs().max().item(), flush=True)
    newp = torch.cat([torch.prod(1 - r, dim=0, keepdim=True), r], 0).clamp(1e-7, 1 - 1e-7)
    ref = F.softmax(torch.log(newp / (1 - newp)), dim=0)
    print("aggregate err", (prob - ref).abs().max().item(), flush=True)
    ag = ops.aggregate_wbg(r, keep_bg=True)
    print("aggregate_wbg err", (ag - ref).abs().max().item(), flush=True)
    agh = ops.aggregate_wbg(r, keep_bg=False, hard=True)
    refh = F.softmax(torch.log(newp / (1 - newp)) * 1000, dim=0)[1:]
    print("aggregate_wbg hard err", (agh - refh).abs().max().item(), flush=True)
    # argmax
    T = 4
    pr = torch.rand(k + 1, T, 1, 48, 64, device=dev)
    mp = torch.zeros((T, 1, 48, 64), dtype=torch.uint8, device=dev)
    mo = torch.zeros((T, 44, 60), dtype=torch.uint8, device=dev)
    ops.argmax_unpad(pr, (2, 2, 2, 2), 44, 60, mp, mo)
    ref = torch.argmax(pr, 0).to(torch.uint8)
    print("argmax exact", bool((mp == ref).all()), bool((mo == ref[:, 0, 2:-2, 2:-2]).all()), flush=True)
    p2 = ops.pad2d(pr, (3, 4, 1, 2))
    print("pad exact", bool((p2 == F.pad(pr, (3, 4, 1, 2))).all()), flush=True)


run(elementwise)


# ------------------------------------------------------------------ exact memory read
def memread(K, T, h, w, top_k, algo):
    hw = h * w
    slots = T * hw
    mk = torch.randn(K, 128, T, h, w, device=dev)
    mv = torch.randn(K, 512, T, h, w, device=dev)
    qk = torch.randn(1, 128, h, w, device=dev)
    cap = slots + 100
    bk = torch.zeros((K, cap, 128), device=dev)
    bv = torch.zeros((K, cap, 512), device=dev)
    ops.bank_from_nchw(mk, mv, bk, bv)
    print("bank_from_nchw exact", bool((bk[:, :slots] == mk.reshape(K, 128, slots).transpose(1, 2)).all()),
          bool((bv[:, :slots] == mv.reshape(K, 512, slots).transpose(1, 2)).all()), flush=True)
    qpm = qk.reshape(128, hw).t().contiguous()
    out = torch.zeros((K, hw, 512), device=dev)
    t0 = time.time()
    out, idx, val = ops.memory_read(bk, bv, slots, qpm, top_k, out, algo=algo, want_topk=True)
    torch.cuda.synchronize()
    _lib.poll_kernel_error()
    # reference (the reference's own formulation, fp64)
    mi = mk.reshape(K, 128, slots).transpose(1, 2).double()
    qi = (qk.reshape(1, 128, hw) / (128 ** 0.5)).double().expand(K, -1, -1)
    aff = torch.bmm(mi, qi)
    vals, ind = torch.topk(aff, top_k, dim=1)
    xe = torch.exp(vals - vals[:, :1])
    xe = xe / xe.sum(1, keepdim=True)
    aff.zero_().scatter_(1, ind, xe)
    ref = torch.bmm(mv.reshape(K, 512, slots).double(), aff).transpose(1, 2)  # K, hw, 512
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    same = (idx.long().transpose(1, 2).sort(1)[0] == ind.sort(1)[0]).all(1)
    print(f"memread K={K} T={T} hw={hw} k={top_k} algo={algo}: readout rel err {err:.3e}; index sets equal in "
          f"{int(same.sum())}/{same.numel()} columns; score err {(val.double().transpose(1,2) - vals).abs().max().item():.3e}", flush=True)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    ws = torch.empty(ops.memory_read_workspace_bytes(K, slots, hw, top_k), dtype=torch.uint8, device=dev)
    for _ in range(2):
        ops.memory_read(bk, bv, slots, qpm, top_k, out, workspace=ws, algo=algo)
    e0.record()
    for _ in range(5):
        ops.memory_read(bk, bv, slots, qpm, top_k, out, workspace=ws, algo=algo)
    e1.record()
    torch.cuda.synchronize()
    print(f"   time {e0.elapsed_time(e1)/5*1e3:.1f} us", flush=True)


run(memread, 1, 3, 30, 54, 20, ops.MEMREAD_EXACT_SIMT)
run(memread, 2, 5, 30, 54, 50, ops.MEMREAD_EXACT_SIMT)
run(memread, 1, 20, 30, 54, 20, ops.MEMREAD_EXACT_SIMT)


def attention():
    h16, w16 = 30, 54
    hw = h16 * w16
    mk = torch.randn(1, 128, 1, h16, w16, device=dev)
    qk = torch.randn(1, 128, h16, w16, device=dev)
    pos = torch.rand(1, 1, h16 * 16, w16 * 16, device=dev)
    neg = torch.rand(1, 1, h16 * 16, w16 * 16, device=dev)
    out = ops.attention_map(mk.reshape(128, hw).t().contiguous(), qk.reshape(128, hw).t().contiguous(), h16, w16, pos, neg)
    torch.cuda.synchronize()
    m = mk.reshape(1, 128, hw).transpose(1, 2)
    q = qk.reshape(1, 128, hw) / (128 ** 0.5)
    Wm = F.softmax(torch.bmm(m, q), dim=1)
    pm = F.interpolate(pos, size=(h16, w16), mode="area").view(1, 1, hw) @ Wm
    nm = F.interpolate(neg, size=(h16, w16), mode="area").view(1, 1, hw) @ Wm
    ref = F.interpolate(torch.cat([pm, nm], 1).reshape(1, 2, h16, w16), mode="bilinear", size=(h16 * 16, w16 * 16), align_corners=False)
    print("attention_map err", (out - ref).abs().max().item(), "scale", ref.abs().max().item(), flush=True)


run(attention)
print("launches", l.mivos_launch_count(), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(RES, open("gpurun_out/probe1.json", "w"), indent=1)
