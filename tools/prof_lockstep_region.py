"""Launch list of a steady-state region of the unit bench.py times: one lock-step lane of 4 cfg-2 clips, eager
launches (MIVOS_GRAPH=0: the same kernels in the same order as the replayed graph), frames 50..54 of a 60-frame
clip (memory bank of 11 frames = the mean bank of the 101-frame clip; one memorize-to-bank frame, one batched
query pass per clip).  Only that region is profiled:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv \\
      python tools/prof_lockstep_region.py
  python tools/ncu_summary.py launches L.csv"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MIVOS_GRAPH"] = "0"
os.environ.setdefault("MIVOS_ACT_DTYPE", "fp16")

import torch  # noqa: E402

torch.set_grad_enabled(False)
import mivos_b200  # noqa: E402
from mivos_b200 import _lib, synth  # noqa: E402

dev = torch.device("cuda:0")
net = mivos_b200.PropagationNetwork(top_k=20)
net.load_state_dict(synth.make_prop_state_dict())
net = net.to(dev)
T, C, FIRST, LAST = 60, 4, 50, 55
clips = [synth.synthetic_clip(T, 480, 854, 1, seed=1234 + c) for c in range(C)]
cores = [mivos_b200.InferenceCore(net, None, im, 1, mem_profile=0, mem_freq=5, device="cuda:0") for im, _ in clips]
state = {"n": 0}


def step_cb():
    state["n"] += 1
    if state["n"] == FIRST:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    elif state["n"] == LAST:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


mivos_b200.LockstepSession(cores).interact([m for _, m in clips], 0, step_cb=step_cb)
torch.cuda.synchronize()
_lib.poll_kernel_error()
print(f"profiled lock-step frames {FIRST}..{LAST - 1} of {T} ({C} clips); library launches in the process:", int(_lib.load().mivos_launch_count()))
