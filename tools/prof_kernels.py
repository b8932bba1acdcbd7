"""Small driver for ncu captures: a handful of launches of the hot kernels at cfg-2 shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mivos_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "memread"):
    K, T, hw, k = 1, 20, 1620, 20
    slots = T * hw
    bk = torch.randn((K, slots, 128), device=dev)
    bv = torch.randn((K, slots, 512), device=dev)
    qk = torch.randn((hw, 128), device=dev)
    ws = torch.empty(ops.memory_read_workspace_bytes(K, slots, hw, k), dtype=torch.uint8, device=dev)
    out = torch.zeros((K, hw, 512), device=dev)
    for _ in range(4):
        ops.memory_read(bk, bv, slots, qk, k, out, workspace=ws, algo=ops.MEMREAD_TCGEN05)
    torch.cuda.synchronize()
if which in ("all", "conv", "expand", "expand4"):
    def conv(n, h, w, cin, cout, ks, res=False):
        dt = torch.float16 if os.environ.get("MIVOS_ACT_DTYPE", "fp16") == "fp16" else torch.float32
        x = torch.randn(n, h + 2, w + 2, cin, device=dev).to(dt)
        wt = torch.randn(cout, cin, ks, ks, device=dev) / (cin * ks * ks) ** 0.5
        pc = ops.pack_conv(wt, torch.zeros(cout, device=dev), device=dev, dtype=dt)
        out = torch.zeros((n, h + 2, w + 2, pc.cout_pad), device=dev, dtype=dt)
        r = torch.randn(n, h + 2, w + 2, pc.cout_pad, device=dev).to(dt) if res else None
        for _ in range(3):
            ops.conv_gemm(x, pc, n, h, w, out, relu=True, round_tf32=True, residual=r)
    if which == "expand4":
        conv(4, 120, 216, 64, 256, 1, res=True)   # the same at the batch of a lock-step step (working set > L2)
        torch.cuda.synchronize()
        print("done")
        sys.exit(0)
    if which == "expand":
        conv(1, 120, 216, 64, 256, 1, res=True)   # bottleneck conv3 + residual: output-bound
        torch.cuda.synchronize()
        print("done")
        sys.exit(0)
    conv(1, 120, 216, 256, 256, 3)   # decoder up_8_4 (largest layers)
    conv(1, 30, 54, 1024, 256, 1)    # layer3 bottleneck 1x1
    conv(1, 30, 54, 256, 256, 3)     # layer3 bottleneck 3x3
    conv(8, 30, 54, 256, 256, 3)     # the same in the batched query pass
    torch.cuda.synchronize()
print("done")
