#!/usr/bin/env python
"""Small runs of the round-2 host paths for compute-sanitizer (memcheck / racecheck / synccheck):
  * LockstepSession of 2 clips x 2 objects (batched query sets in the memory read, pointer-table graph replay,
    single bank_write, batched stems / upsamplers),
  * one InferenceCore.interact on a MIDDLE frame (forward and backward pass as two concurrent lanes),
both with fp16 maps (TMA epilogue, 8-warp tiles, tap reuse) at 96x128 so the whole thing finishes in minutes
under the tools.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_lockstep.py [fp16|tf32]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
act = sys.argv[1] if len(sys.argv) > 1 else "fp16"
os.environ["MIVOS_ACT_DTYPE"] = act

import torch  # noqa: E402

torch.set_grad_enabled(False)
import mivos_b200  # noqa: E402
from mivos_b200 import _lib, synth  # noqa: E402

dev = torch.device("cuda:0")
net = mivos_b200.PropagationNetwork(top_k=20)
net.load_state_dict(synth.make_prop_state_dict())
net = net.to(dev)
K, T, C = 2, 7, 2
clips = [synth.synthetic_clip(T, 96, 128, K, seed=11 + c) for c in range(C)]
cores = [mivos_b200.InferenceCore(net, None, im, K, mem_freq=2, device="cuda:0") for im, _ in clips]
out = mivos_b200.LockstepSession(cores).interact([m for _, m in clips], 0)
torch.cuda.synchronize()
_lib.poll_kernel_error()
print("lock-step:", [o.shape for o in out], "launches", int(_lib.load().mivos_launch_count()))
images, mask = synth.synthetic_clip(T, 96, 128, K, seed=5)
core = mivos_b200.InferenceCore(net, None, images, K, mem_freq=2, device="cuda:0")
m = core.interact(mask, 3)
torch.cuda.synchronize()
_lib.poll_kernel_error()
print("middle-frame interaction (two lanes):", m.shape, "launches", int(_lib.load().mivos_launch_count()))
