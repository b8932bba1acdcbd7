"""Times the S2M network at DAVIS 480p on cuda:0: K objects of one interaction as one
batch, CUDA events, after warm-up.  Usage: python tools/s2m_time.py [fp16|tf32]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mivos_b200  # noqa: E402
from mivos_b200 import _lib, synth  # noqa: E402

torch.set_grad_enabled(False)
act = sys.argv[1] if len(sys.argv) > 1 else "fp16"
dev = torch.device("cuda:0")
net = mivos_b200.S2MNetwork(act_dtype=torch.float16 if act == "fp16" else torch.float32)
net.load_state_dict(synth.make_s2m_state_dict())
net = net.to(dev)
for k in (1, 3):
    x = torch.randn((k, 6, 480, 864), device=dev)
    for _ in range(3):
        net.forward_sigmoid(x)
    torch.cuda.synchronize()
    l0 = int(_lib.load().mivos_launch_count())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        net.forward_sigmoid(x)
    e1.record()
    torch.cuda.synchronize()
    _lib.poll_kernel_error()
    launches = (int(_lib.load().mivos_launch_count()) - l0) // reps
    ms = e0.elapsed_time(e1) / reps
    # 6-channel ResNet-50 at output stride 16 + ASPP + head: FLOPs counted from the layer table
    print(f"s2m {act} 480x864 K={k}: {ms:.3f} ms per interaction ({ms / k:.3f} ms/object), {launches} launches, "
          f"workspace {net.engine().ws.bytes() / 2**20:.0f} MiB", flush=True)
