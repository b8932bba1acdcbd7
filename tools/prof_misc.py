"""Launches of the kernels outside the propagation loop's top list, for one `ncu --set full` capture:
S2M (gather_dilated, halo_avgpool_broadcast, upsample_bilinear), the DAVIS overlay, and a split-K convolution
(its partial-sum epilogue kernel).  Usage:
  ncu --set full -k regex:"gather_dilated|halo_avgpool|upsample_bilinear|overlay|splitk_epilogue" -c 8 python tools/prof_misc.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mivos_b200  # noqa: E402
from mivos_b200 import _lib, ops, synth  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
net = mivos_b200.S2MNetwork(act_dtype=torch.float16)
net.load_state_dict(synth.make_s2m_state_dict())
net = net.to(dev)
x = torch.randn((1, 6, 480, 864), device=dev)
for _ in range(2):
    net.forward_sigmoid(x)
torch.cuda.synchronize()
img = torch.randint(0, 255, (8, 480, 854, 3), dtype=torch.uint8, device=dev)
msk = torch.randint(0, 3, (8, 480, 854), dtype=torch.uint8, device=dev)
for _ in range(2):
    ops.overlay_davis(img, msk, 0.5)
torch.cuda.synchronize()
# a 1/16-resolution layer of the sequential step at batch 1: few row tiles, long K -> the split-K plan
dt = torch.float16
n, h, w, cin, cout = 1, 30, 54, 1024, 512
xh = torch.randn(n, h + 2, w + 2, cin, device=dev).to(dt)
wt = torch.randn(cout, cin, 3, 3, device=dev) / (cin * 9) ** 0.5
pc = ops.pack_conv(wt, torch.zeros(cout, device=dev), device=dev, dtype=dt)
out = torch.zeros((n, h + 2, w + 2, pc.cout_pad), device=dev, dtype=dt)
ws = ops.split_k_workspace(dev)
for _ in range(3):
    ops.conv_gemm(xh, pc, n, h, w, out, relu=True, round_tf32=True, splitk_ws=ws)
torch.cuda.synchronize()
_lib.poll_kernel_error()
print("done")
