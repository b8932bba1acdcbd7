#!/usr/bin/env python
"""bench.py — propagated frames/sec of the mask-propagation hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W                (ours; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W   (reference algorithm on host cores)

Workload (BASELINE configs[1]): DAVIS-2017-val-shaped synthetic clip,
480x854 (padded to 480x864), 1 object, mem_freq 5, top-k 20, seeded random weights in the
reference's checkpoint format.  ONE STEP = one `InferenceCore.interact(mask, 0)` over a T-frame
clip = T-1 propagated frames (memory bank grows 1 -> (T-2)//5+2 frames, 20+1 at T=101), for each
of the clips a GPU propagates at a time: `--clips-per-gpu` concurrent lanes (default 2: own network
object, own CUDA stream, one Python thread each) x `--lockstep` clips per lane (default 4) that
advance together as ONE batch through every convolution (mivos_b200.LockstepSession: the per-frame
chain of one clip is ~70 short dependent kernels whose 1/16- and 1/8-resolution layers have 14-54
row tiles for 148 SMs).  `--clips-per-gpu 1 --lockstep 1` gives the single-clip number.
  value : frames/s with the clip resident in HBM when the timed region starts (mem_profile=0)
  e2e   : same call with the clip in PINNED HOST memory (mem_profile=1): every frame is copied
          H2D inside the timed region and the u8 masks are copied D2H at the end.
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
Every step reads a 101-frame clip (503 MB) plus >200 MB of weights: inputs exceed the 126 MB L2.
Multi-GPU: clips shard across ranks (clips_per_gpu clips per rank per step, weak scaling, no data-path
collective); NCCL only for the barrier and the max/sum reductions of the timings.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, K_OBJ, MEM_FREQ, TOP_K = 480, 854, 1, 5, 20
METRIC = "propagated frames/sec, 480p, 1 object (mask propagation)"
DEFAULT_ACT = "fp16"
ACT_DTYPE = torch.float32  # set from --act in main()


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe).  One looping
    `nvidia-smi -lms` child, 500 ms period: every NVML query briefly stalls concurrent launches
    (measured: an in-process pynvml thread at 200 ms cost 15 % of `value`; the e2e region, which is
    not sampled, shows the unperturbed rate)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("MIVOS_BENCH_SMI_MS", "500")], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _best_cpu_threads(psd, frame):
    """The reference runs `torch` CPU ops on however many threads PyTorch is given.  On a many-core
    host oversubscription can make it SLOWER (measured: 128 threads are >10x slower than 16 on the
    GPU box), so give the CPU side its best case: time one query-encoder call (73 GFLOP of convs)
    at a few thread counts and keep the fastest.  MIVOS_CPU_THREADS overrides."""
    from oracle import stm_oracle as O
    env = os.environ.get("MIVOS_CPU_THREADS")
    if env:
        torch.set_num_threads(int(env))
        return int(env)
    ncpu = os.cpu_count() or 1
    # (PyTorch CPU convolutions stop scaling long before 64 threads; the sweep itself is bounded)
    cands = sorted({c for c in (min(ncpu, 64), 32, 16, 8) if 1 <= c <= ncpu}, reverse=True)
    best, best_t = cands[-1], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        O.get_query_values(psd, frame)  # warm the thread pool
        t0 = time.perf_counter()
        O.get_query_values(psd, frame)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def _dist():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's algorithm on the host cores: the CPU oracle port (oracle/stm_oracle.py — the
    reference itself is Python and cannot travel to the GPU box), all host threads.
    Each step is a BOUNDED sample of the same workload: interact() on the first `ref_frames` frames
    of the clip (same 480p shapes, bank grows from 1 frame)."""
    rank, world, _ = _dist()
    if rank != 0:
        return
    from oracle import stm_oracle as O
    from mivos_b200 import synth
    torch.set_grad_enabled(False)
    psd = synth.make_prop_state_dict()
    images, mask = synth.synthetic_clip(args.ref_frames, H, W, K_OBJ, seed=1234)
    cores = _best_cpu_threads(psd, O.pad_divide_by(images[:, 0], 16)[0])
    frames = args.ref_frames - 1
    times = []
    for i in range(args.warmup + args.steps):
        core = O.OracleInferenceCore(psd, None, images, K_OBJ, mem_freq=MEM_FREQ, top_k=TOP_K)
        t0 = time.perf_counter()
        core.interact(mask, 0)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    fps = frames * len(times) / total
    sample = f"interact() on the first {args.ref_frames} frames (={frames} propagated) of the 480p clip per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 480p, 1 object, mem_freq 5, top-k 20", "frames_per_step": frames},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "host_cores": os.cpu_count(), "kind": "port", "sample": sample,
                         "torch": torch.__version__},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------- our arm
def _bracketed_pass(mivos_b200, ops, net, clips, masks, dev):
    """One extra EAGER interact() of `clips` (one clip, or several in lock-step) with CUDA events around
    every convolution and memory-read call on the launching stream.  Returns (records, GPU ms of the pass).
    Graph replay is switched off for the pass (per-launch events need eager launches: same kernels, same
    order) and restored afterwards."""
    rec = {"conv": [], "memread": []}
    orig_conv, orig_mr = ops.conv_gemm, ops.memory_read

    def conv_prof(x, pc, n, h, w, out, **kw):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = orig_conv(x, pc, n, h, w, out, **kw)
        b.record()
        rec["conv"].append((a, b, 2.0 * n * h * w * pc.ksize * pc.ksize * pc.cin * pc.cout))
        return r

    def mr_prof(bank_k, bank_v, slots, qk, top_k, out, **kw):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = orig_mr(bank_k, bank_v, slots, qk, top_k, out, **kw)
        b.record()
        if kw.get("dyn_slots") is not None:  # lock-step step: `slots` is the capacity, the live count is on the device
            slots = int(kw["dyn_slots"][0])
        kk, hw = bank_k.shape[0], qk.shape[0]
        rec["memread"].append((a, b, 2.0 * 128 * slots * hw * kk + 2.0 * top_k * 512 * hw * kk,
                               4.0 * (kk * slots * 128 + kk * top_k * hw * 512 + hw * 128 + kk * hw * 512)))
        return r

    lock_steps = list(net.engine().__dict__.get("_lock_steps", {}).values())
    saved_flags = [st.use_graph for st in lock_steps]
    saved_env = os.environ.get("MIVOS_GRAPH")
    ops.conv_gemm, ops.memory_read = conv_prof, mr_prof
    os.environ["MIVOS_GRAPH"] = "0"
    for st in lock_steps:
        st.use_graph = False
    try:
        cores = [mivos_b200.InferenceCore(net, None, im, K_OBJ, mem_profile=0, mem_freq=MEM_FREQ, device=dev) for im in clips]
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        if len(cores) == 1:
            cores[0].interact(masks[0], 0)
        else:
            mivos_b200.LockstepSession(cores).interact(masks, 0)
        pe1.record()
        torch.cuda.synchronize()
        prof_ms = pe0.elapsed_time(pe1)  # GPU time of this pass: the denominator of the shares
    finally:
        ops.conv_gemm, ops.memory_read = orig_conv, orig_mr
        if saved_env is None:
            os.environ.pop("MIVOS_GRAPH", None)
        else:
            os.environ["MIVOS_GRAPH"] = saved_env
        for st, flag in zip(lock_steps, saved_flags):
            st.use_graph = flag
    return rec, prof_ms


def _roofline_dicts(rec, prof_ms, peaks, peak_src, fp16, what):
    conv_ms = sum(a.elapsed_time(b) for a, b, _ in rec["conv"])
    conv_fl = sum(f for _, _, f in rec["conv"])
    tf32_peak = peaks["bf16_tflops_sustained"] / 2.0
    conv_peak = peaks["bf16_tflops_sustained"] if fp16 else tf32_peak
    ach = conv_fl / (conv_ms / 1e3) / 1e12
    roof = {"kernel": f"conv_gemm_persistent_kernel (tcgen05 kind::{'f16' if fp16 else 'tf32'} implicit GEMM, all conv layers of the step)",
            "bound": "tensor", "achieved": ach, "peak": conv_peak, "unit": "TFLOP/s", "frac": ach / conv_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the largest layer
            # (3x3 256->256 @120x216) from the `ncu --set full` capture summarised in
            # profiles/r01_ncu_full_summaries_fp16.txt (fp16) / r01_ncu_full_summaries.txt (tf32);
            # algorithmic bytes of that launch: 13.6 + 13.6 MB maps + 1.2 MB weights (fp16)
            "traffic": (14884608 if fp16 else 53218304), "traffic_launch": "conv 3x3 256->256 @120x216 n=1",
            "launches": len(rec["conv"]), "avg_launch_us": 1e3 * conv_ms / max(1, len(rec["conv"])),
            "share_of_step": conv_ms / prof_ms, "measured_on": what,
            "peak_source": (f"{peak_src}: sustained dense bf16 {peaks['bf16_tflops_sustained']:.0f} (kind::f16 issues at the bf16 rate)" if fp16 else
                            f"{peak_src}: TF32 dense = sustained bf16 {peaks['bf16_tflops_sustained']:.0f} / 2 (kind::tf32 issues at half the bf16 rate)"),
            "note": "event-bracketed launches are serialised: durations include launch gaps, so this is a lower bound; "
                    "share_of_step = bracketed time / GPU time of the same eager pass"}
    mr_ms = sum(a.elapsed_time(b) for a, b, _, _ in rec["memread"])
    mr_fl = sum(f for _, _, f, _ in rec["memread"])
    mr_by = sum(by for _, _, _, by in rec["memread"])
    roof_mr = {"kernel": "memory_read (prep + memread_tc_kernel + select)", "bound": "tensor",
               "achieved": mr_fl / (mr_ms / 1e3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
               "frac": mr_fl / (mr_ms / 1e3) / 1e12 / tf32_peak, "hbm_gbs": mr_by / (mr_ms / 1e3) / 1e9,
               "hbm_frac": mr_by / (mr_ms / 1e3) / 1e9 / peaks["hbm_gbs"], "launches": len(rec["memread"]),
               "avg_call_us": 1e3 * mr_ms / max(1, len(rec["memread"])), "share_of_step": mr_ms / prof_ms, "measured_on": what}
    return roof, roof_mr


def run_ours(args):
    rank, world, local = _dist()
    torch.set_grad_enabled(False)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    import mivos_b200
    from mivos_b200 import _lib, ops, sharding, synth

    import threading as _th

    C = max(1, args.clips_per_gpu)
    L = max(1, args.lockstep)  # clips a lane advances in lock-step as one batch (mivos_b200/lockstep.py); 1 = off
    sd = synth.make_prop_state_dict()
    T = args.frames
    frames = T - 1
    nh, nw = 480, 864

    class Lane:
        """One clip slot of this GPU: own network object (own packed weights, workspaces and captured
        graphs), own CUDA stream, own clip.  Lanes run concurrently from Python threads; their kernels
        interleave on the device and fill the SMs the latency-bound per-frame chain leaves idle."""

        def __init__(self, i):
            self.net = mivos_b200.PropagationNetwork(top_k=TOP_K, act_dtype=ACT_DTYPE)
            self.net.load_state_dict(sd)
            self.net = self.net.to(dev)
            self.stream = torch.cuda.Stream(device=dev)
            mine = sharding.clips_of_rank(world * C * L, rank, world)  # clip c -> rank c % world
            self.clips = mine[i * L:(i + 1) * L]
            self.clip = self.clips[0]
            data = [synth.synthetic_clip(T, H, W, K_OBJ, seed=1234 + c) for c in self.clips]
            self.images_l, self.masks_l = [d[0] for d in data], [d[1] for d in data]
            self.images, self.mask = self.images_l[0], self.masks_l[0]
            self.checksums = [0] * L
            self.checksum = 0
            self.results = []
            self.step_wall = []

        def run(self, cores, nsteps):
            """`nsteps` steps over this lane's L sessions, reused from step to step: InferenceCore.reset()
            returns a session to its freshly-constructed state (query cache dropped, certain memories
            forgotten), so every step recomputes everything; what is NOT repeated is the construction —
            clip upload / pinning and buffer allocation — which the metric excludes.
            Keeping one set of sessions alive instead of one per step bounds device memory at
            C*L sessions (3.5 GB each at 101 frames: clip, probabilities, the 105-frame query cache)."""
            torch.cuda.set_device(dev)
            with torch.cuda.stream(self.stream):
                for _ in range(nsteps):
                    # interact() returns the host u8 masks of the clip: the D2H read of the step's
                    # result is inside the timed region, the checksum over them is not
                    t0 = time.perf_counter()
                    for c in cores:
                        c.reset()
                    if L == 1:
                        self.results.append([cores[0].interact(self.mask, 0)])
                    else:
                        self.results.append(mivos_b200.LockstepSession(cores).interact(self.masks_l, 0))
                    self.step_wall.append(time.perf_counter() - t0)  # interact() ends with a stream sync

        def take_checksum(self):
            for per_clip in self.results:
                for j, m in enumerate(per_clip):
                    self.checksums[j] += int(m.sum(dtype="int64"))
            self.checksum = sum(self.checksums)
            self.results = []

    lanes = [Lane(i) for i in range(C)]
    net, images, mask = lanes[0].net, lanes[0].images, lanes[0].mask

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_region(mem_profile, nsteps, warm):
        cores = [[mivos_b200.InferenceCore(ln.net, None, ln.images_l[j], K_OBJ, mem_profile=mem_profile, mem_freq=MEM_FREQ,
                                           device=dev) for j in range(L)] for ln in lanes]
        for ln, cs in zip(lanes, cores):  # warm-up lane by lane (graph capture is single-threaded)
            ln.run(cs, warm)
            ln.results = []
            ln.step_wall = []
        barrier()
        l0 = _lib.load().mivos_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream(dev)
        t0 = time.perf_counter()
        e0.record(main)
        for ln in lanes:
            ln.stream.wait_event(e0)
        threads = [_th.Thread(target=ln.run, args=(cs, nsteps)) for ln, cs in zip(lanes, cores)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        for ln in lanes:
            main.wait_stream(ln.stream)
        e1.record(main)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        launches = _lib.load().mivos_launch_count() - l0
        for ln in lanes:
            ln.take_checksum()
        step_ms = [[round(1e3 * x, 2) for x in ln.step_wall] for ln in lanes]
        per_clip = sharding.gather_clip_results([(c, ln.checksums[j]) for ln in lanes for j, c in enumerate(ln.clips)],
                                                world * C * L)
        return sharding.max_over_ranks(ms, dev), launches, sum(per_clip), wall, step_ms

    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, launches, checksum, wall_dev, steps_dev = timed_region(0, args.steps, args.warmup)
    clocks = sampler.stop()
    ms_e2e, _, checksum2, wall_e2e, steps_e2e = timed_region(1, args.steps, max(1, args.warmup // 3))
    value = world * C * L * frames * args.steps / (ms_dev / 1e3)
    e2e = world * C * L * frames * args.steps / (ms_e2e / 1e3)
    h2d = world * C * L * (T * 3 * nh * nw * 4 + (K_OBJ + 1) * H * W * 4)  # whole job, like `value`
    d2h = world * C * L * T * H * W

    # ---------------- roofline of the dominant kernel (conv implicit GEMM) + the memory read,
    # measured live with CUDA events around each launch on the launching stream (rank 0)
    roof, roof_mr, cpu, roof_extra = None, None, None, {}
    if rank == 0:
        peaks, peak_src = _peaks()
        fp16 = ACT_DTYPE == torch.float16
        # (a) one clip, eager — the pass every earlier profile of this repo refers to
        try:
            rec, prof_ms = _bracketed_pass(mivos_b200, ops, net, [images], [mask], dev)
            roof, roof_mr = _roofline_dicts(rec, prof_ms, peaks, peak_src, fp16, "one clip, eager")
        except Exception as e:  # never lose the timed numbers to a failure of the explanatory pass
            roof_extra = {"roofline_error": f"{type(e).__name__}: {e}"}
        if L > 1 and roof is not None:
            # (b) the unit of the timed workload: one lane = L clips in lock-step (C*K maps per conv launch)
            try:
                rec, prof_ms = _bracketed_pass(mivos_b200, ops, net, lanes[0].images_l, lanes[0].masks_l, dev)
                r2, m2 = _roofline_dicts(rec, prof_ms, peaks, peak_src, fp16, f"one lane: {L} clips in lock-step, eager")
                roof_extra = {"roofline_single_clip": roof, "roofline_memory_read_single_clip": roof_mr}
                roof, roof_mr = r2, m2
            except Exception as e:  # keep the bench line: (a) stands, the failure is reported
                roof_extra = {"roofline_lockstep_error": f"{type(e).__name__}: {e}"}

        # ---------------- CPU baseline: the oracle port on this box's host cores, bounded sample
        if world == 1 and not args.skip_cpu_baseline:
            from oracle import stm_oracle as O
            psd = synth.make_prop_state_dict()
            ncores = _best_cpu_threads(psd, O.pad_divide_by(images[:, 0], 16)[0])
            n = args.ref_frames
            oc = O.OracleInferenceCore(psd, None, images[:, :n].contiguous(), K_OBJ, mem_freq=MEM_FREQ, top_k=TOP_K)
            t0 = time.perf_counter()
            om = oc.interact(mask, 0)
            dt = time.perf_counter() - t0
            core = mivos_b200.InferenceCore(net, None, images[:, :n].contiguous(), K_OBJ, mem_freq=MEM_FREQ, device=dev)
            gm = core.interact(mask, 0)
            cpu = {"value": (n - 1) / dt, "unit": "frames/s", "cores": ncores, "kind": "port",
                   "sample": f"oracle interact() on the first {n} frames ({n-1} propagated) of the same clip, {dt:.1f} s, "
                             f"{ncores} torch threads (fastest of a sweep up to {os.cpu_count()} host cores)",
                   "mask_mismatch_vs_gpu": float((om != gm).mean()), "torch": torch.__version__}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16" if ACT_DTYPE == torch.float16 else "tf32", "data": "synthetic",
            "config": {"workload": f"cfg2: DAVIS-shaped 480p ({H}x{W} -> 480x864), 1 object, {T}-frame clip/rank/step, mem_freq 5, "
                                   f"bank 1->{(T - 2) // MEM_FREQ + 2} frames, top-k 20", "frames_per_step": frames * C * L * world,
                       "clips_per_step": world * C * L, "clips_per_gpu": C * L, "lockstep": L,
                       "parallelism": f"clip-sharded: {world} GPU(s) x {C} concurrent lane(s) per GPU (one CUDA stream + one thread per lane)"
                                      + (f" x {L} clips advanced in lock-step as one batch per lane" if L > 1 else ""), "l2": "inputs larger than L2 (503 MB clip + 215 MB weights per step)"},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "path": "InferenceCore(mem_profile=1).interact(): pinned host clip, per-frame H2D, masks D2H"},
            "gpu_launches": launches, "launch_mode": "cuda-graph replay per frame" if os.environ.get("MIVOS_GRAPH", "1") != "0" else "eager",
            "clocks": clocks, "roofline": roof, "roofline_memory_read": roof_mr, **roof_extra,
            "cpu_baseline": cpu, "mask_checksum": [checksum, checksum2], "wall_s": [wall_dev, wall_e2e],
            "interact_wall_ms": {"resident": steps_dev, "e2e": steps_e2e},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=101, help="clip length per step (cfg-2: 101)")
    ap.add_argument("--ref-frames", type=int, default=4, help="frames of the bounded CPU sample")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--act", default=os.environ.get("MIVOS_ACT_DTYPE", DEFAULT_ACT), choices=["tf32", "fp16"],
                    help="convolution operand / activation type (fp16 = the reference GUI's autocast precision)")
    ap.add_argument("--clips-per-gpu", type=int, default=int(os.environ.get("MIVOS_CLIPS_PER_GPU", "2")),
                    help="concurrent lanes per GPU (own network object, CUDA stream and Python thread each); a lane advances "
                         "--lockstep clips together, so a GPU propagates clips-per-gpu x lockstep clips at a time")
    ap.add_argument("--lockstep", type=int, default=int(os.environ.get("MIVOS_LOCKSTEP", "4")),
                    help="clips each lane advances in lock-step as ONE batch through the conv layers "
                         "(mivos_b200.LockstepSession); 1 = off.  Measured on B200 (profiles/r01b_bench_lockstep_*.json): "
                         "2 lanes x 4 clips 907 frames/s, 1 x 8 855, 2 x 2 792, 1 x 4 687, 2 x 1 690")
    args = ap.parse_args()
    global ACT_DTYPE
    ACT_DTYPE = torch.float16 if args.act == "fp16" else torch.float32
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
