"""Frames/s of cfg-2-shaped propagation (480p, 1 object, mem_freq 5, top-k 20) on cuda:0 as a function of
the number of clips advanced in lock-step (mivos_b200.LockstepSession) — the A/B behind bench.py's
--lockstep.  CUDA-event timing of whole interact() calls after one warm-up call (graph capture).
Usage: python tools/lockstep_sweep.py [frames=41] [L ...=1 2 4 8]
Other BASELINE configs through the environment: SWEEP_SIZE=720x1280 SWEEP_OBJECTS=5 SWEEP_TOPK=50
(cfg-5 shape), SWEEP_OBJECTS=3 SWEEP_TOPK=50 (cfg-3 shape)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mivos_b200  # noqa: E402
from mivos_b200 import _lib, synth  # noqa: E402

torch.set_grad_enabled(False)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 41
Ls = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
dev = torch.device("cuda:0")
SH, SW = (int(v) for v in os.environ.get("SWEEP_SIZE", "480x854").split("x"))
KOBJ, TOPK = int(os.environ.get("SWEEP_OBJECTS", "1")), int(os.environ.get("SWEEP_TOPK", "20"))
net = mivos_b200.PropagationNetwork(top_k=TOPK)
net.load_state_dict(synth.make_prop_state_dict())
net = net.to(dev)
clips = [synth.synthetic_clip(T, SH, SW, KOBJ, seed=100 + i) for i in range(max(Ls))]


def make(L):
    return [mivos_b200.InferenceCore(net, None, clips[i][0], KOBJ, mem_freq=5, device="cuda:0") for i in range(L)]


def run(L, cores):
    masks = [clips[i][1] for i in range(L)]
    if L == 1:
        return [cores[0].interact(masks[0], 0)]
    return mivos_b200.LockstepSession(cores).interact(masks, 0)


for L in Ls:
    run(L, make(L))  # warm-up: workspaces, graph capture
    reps = 2
    fresh = [make(L) for _ in range(reps)]  # sessions are built (clips uploaded) outside the timed region
    torch.cuda.synchronize()
    n0 = int(_lib.load().mivos_launch_count())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for r in range(reps):
        out = run(L, fresh[r])
    e1.record()
    torch.cuda.synchronize()
    _lib.poll_kernel_error()
    ms = e0.elapsed_time(e1) / reps
    frames = L * (T - 1)
    print(f"{SH}x{SW} K={KOBJ} top-k {TOPK} lockstep L={L}: {frames / ms * 1e3:8.1f} frames/s  ({ms / (T - 1):.3f} ms per lock-step frame, "
          f"{(int(_lib.load().mivos_launch_count()) - n0) // reps // (T - 1)} kernels per frame, wall {time.perf_counter() - t0:.2f} s, "
          f"mask sum {sum(int(o.sum()) for o in out)})", flush=True)
