#!/bin/bash
# Round 2, GPU call 10: memory-read emission in PTX + barrier back-off + batched candidate loads (A/B), the joint
# query pass (un-gated test + A/B), sanitizers on the lock-step / two-lane paths, launch list of a steady-state
# lock-step region, ncu --set full of the memory-read kernels on real data, bench lines of cfg3 / cfg4 / cfg5.
set -u
mkdir -p gpurun_out
O=gpurun_out
P=r02c10
echo "== 1. GPU suite"
: > $O/${P}_pytest_gpu.log
for f in tests/test_gpu_*.py; do
  echo "=== $f" >> $O/${P}_pytest_gpu.log
  (timeout 600 python -m pytest $f -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "^$" | tail -40 >> $O/${P}_pytest_gpu.log)
  echo "$f: $(grep -E 'passed|failed|error' $O/${P}_pytest_gpu.log | tail -1)"
done
grep -E "^FAILED|^ERROR|^E  " $O/${P}_pytest_gpu.log | cut -c1-240 | head -30
echo "== 2. memory read A/B (2 lanes x 4 clips, steps 3): emission / back-off / joint query pass"
run_ab() {  # name, env...
  local name=$1; shift
  (env "$@" timeout 300 python bench.py --clips-per-gpu 2 --lockstep 4 --steps 3 --warmup 2 --skip-cpu-baseline --skip-extras --skip-cuda-eager > $O/${P}_ab_$name.json 2> $O/${P}_ab_$name.err)
  python - <<PY
import json
try:
    d = json.load(open("$O/${P}_ab_$name.json"))
    print("%-22s value %.1f e2e %.1f roofline.frac %.3f memread %.1f us launches %d" % ("$name", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_memory_read"]["avg_call_us"], d["gpu_launches"]))
except Exception as e:
    print("$name: failed", e); print(open("$O/${P}_ab_$name.err").read()[-600:])
PY
}
run_ab new X=1
run_ab emit_c_backoff0 MIVOS_MEMREAD_EMIT=c MIVOS_MEMREAD_BACKOFF_NS=0
run_ab joint_query MIVOS_LOCKSTEP_JOINT_QUERY=1
echo "== 3. per-layer table"
(timeout 120 python tools/lockstep_layer_times.py 4 fp16 > $O/${P}_layers_lockstep4.log 2>&1); head -8 $O/${P}_layers_lockstep4.log
echo "== 4. launch list of a steady-state lock-step region (eager launches, all kernels)"
(MIVOS_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2600 --csv --log-file $O/${P}_lockstep_launches.csv \
  python bench.py --clips-per-gpu 1 --lockstep 4 --steps 1 --warmup 1 --skip-cpu-baseline --skip-extras --skip-cuda-eager --skip-roofline > $O/${P}_bench_under_ncu2.log 2>&1)
python tools/ncu_summary.py launches $O/${P}_lockstep_launches.csv > $O/${P}_lockstep_launch_list.txt 2>&1; head -24 $O/${P}_lockstep_launch_list.txt
echo "== 5. ncu --set full: memory-read kernels of a 4-clip lock-step call late in the clip (real data)"
(MIVOS_GRAPH=0 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"memread_(tc|select)" -s 160 -c 4 -o $O/${P}_memread_real \
  python bench.py --clips-per-gpu 1 --lockstep 4 --steps 1 --warmup 1 --skip-cpu-baseline --skip-extras --skip-cuda-eager --skip-roofline > $O/${P}_bench_under_ncu3.log 2>&1)
ls -la $O/${P}_memread_real.ncu-rep
echo "== 6. sanitizers on the lock-step session and the two-lane interaction"
for tool in memcheck racecheck synccheck; do
  (timeout 240 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_lockstep.py fp16 > $O/${P}_sanitizer_${tool}_lockstep.log 2>&1); echo "$tool lockstep rc=$?"; tail -3 $O/${P}_sanitizer_${tool}_lockstep.log
done
echo "== 7. bench lines of the other configurations (ours + reference arm)"
for c in cfg3 cfg4 cfg5; do
  (timeout 900 python bench.py --config $c --steps 3 --warmup 3 --extra-steps 2 > $O/${P}_bench_$c.json 2> $O/${P}_bench_$c.err)
  python - <<PY
import json
try:
    d = json.load(open("$O/${P}_bench_$c.json"))
    e = d.get("reference_cuda_eager") or {}
    print("$c: value %.1f e2e %.1f | single %s / %s | tf32 %s / %s | eager fp32 %s | cpu %s | roofline %.3f | lanes %s" % (
        d["value"], d["e2e"]["value"], d.get("value_single_session"), d.get("e2e_single_session"), d.get("value_tf32"), d.get("e2e_tf32"),
        e.get("fp32", {}).get("value"), d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 3), d["roofline"]["frac"], d["config"].get("name")))
except Exception as ex:
    print("$c failed:", ex); print(open("$O/${P}_bench_$c.err").read()[-800:])
PY
done
for c in cfg3 cfg4; do
  (timeout 600 python bench.py --impl reference --config $c --steps 1 --warmup 1 > $O/${P}_bench_${c}_reference.json 2> $O/${P}_bench_${c}_reference.err); tail -c 300 $O/${P}_bench_${c}_reference.json; echo
done
echo "== done"
