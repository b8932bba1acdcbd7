#!/bin/bash
# Round 2, GPU call 9: recalibrated tile plan (output-bound layers on the 8-warp tiles); whole bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/r02c9_pytest_gpu.log
for f in tests/test_gpu_*.py; do
  echo "=== $f" >> $O/r02c9_pytest_gpu.log
  (timeout 600 python -m pytest $f -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "^$" | tail -40 >> $O/r02c9_pytest_gpu.log)
  echo "$f: $(grep -E 'passed|failed|error' $O/r02c9_pytest_gpu.log | tail -1)"
done
grep -E "^FAILED|^ERROR|^E  " $O/r02c9_pytest_gpu.log | cut -c1-240 | head -30
echo "== 2. per-layer tables"
(timeout 120 python tools/lockstep_layer_times.py 4 fp16 > $O/r02c9_layers_lockstep4.log 2>&1); head -6 $O/r02c9_layers_lockstep4.log; grep "+res" $O/r02c9_layers_lockstep4.log | head -8
(timeout 120 python tools/lockstep_layer_times.py 1 fp16 > $O/r02c9_layers_lockstep1.log 2>&1); head -6 $O/r02c9_layers_lockstep1.log
(timeout 120 python tools/time_phases.py > $O/r02c9_phase_times.log 2>&1); tail -8 $O/r02c9_phase_times.log
echo "== 3. bench A/B"
for cfg in "2 4" "3 4" "1 1"; do
  set -- $cfg
  (timeout 300 python bench.py --clips-per-gpu $1 --lockstep $2 --steps 3 --warmup 2 --skip-cpu-baseline --skip-extras --skip-cuda-eager > $O/r02c9_bench_c$1_l$2.json 2> $O/r02c9_bench_c$1_l$2.err)
  python - <<PY
import json
try:
    d = json.load(open("$O/r02c9_bench_c$1_l$2.json"))
    print("lanes $1 x clips $2: value %.1f e2e %.1f roofline.frac %.3f memread %.1f us launches %d" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_memory_read"]["avg_call_us"], d["gpu_launches"]))
except Exception as e:
    print("lanes $1 x clips $2: failed", e); print(open("$O/r02c9_bench_c$1_l$2.err").read()[-600:])
PY
done
echo "== 4. the whole default bench line + the reference arm (driver's flags)"
(timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02c9_bench_default.json 2> $O/r02c9_bench_default.err); tail -c 400 $O/r02c9_bench_default.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02c9_bench_default.json"))
    e = d.get("reference_cuda_eager") or {}
    print("value %.1f e2e %.1f | single %s / %s | tf32 %s / %s | eager fp32 %s autocast %s | cpu %s | roofline %.3f (tf32 %s) memread %.1f us | clocks %s" % (
        d["value"], d["e2e"]["value"], d.get("value_single_session"), d.get("e2e_single_session"), d.get("value_tf32"), d.get("e2e_tf32"),
        e.get("fp32", {}).get("value"), e.get("autocast_fp16", {}).get("value"), d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 2),
        d["roofline"]["frac"], (d.get("roofline_tf32") or {}).get("frac"), d["roofline_memory_read"]["avg_call_us"], d["clocks"]))
except Exception as ex:
    print("bench line failed:", ex)
PY
(timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02c9_bench_reference.json 2> $O/r02c9_bench_reference.err); python -c "
import json; d=json.load(open('gpurun_out/r02c9_bench_reference.json')); print('reference arm: value %.2f frames/s, ms_per_step %.0f, cores %s, sample: %s' % (d['value'], d['ms_per_step'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample'][:90]))"
echo "== done"
