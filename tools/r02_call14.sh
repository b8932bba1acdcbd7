#!/bin/bash
# Round 2, GPU call 14: the whole GPU suite on the final memory read (sum rule for the exact fallback, pair
# threshold, pre-filtering selection stage, TMEM reads pipelined in the candidate pass), the default bench line and
# the reference arm with the driver's flags, cfg3 / cfg5 lines, launch list of a steady-state lock-step region.
set -u
mkdir -p gpurun_out
O=gpurun_out
P=r02c14
echo "== 1. GPU suite"
: > $O/${P}_pytest_gpu.log
for f in tests/test_gpu_*.py; do
  echo "=== $f" >> $O/${P}_pytest_gpu.log
  (timeout 400 python -m pytest $f -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "^$" | tail -40 >> $O/${P}_pytest_gpu.log)
  echo "$f: $(grep -E 'passed|failed|error' $O/${P}_pytest_gpu.log | tail -1)"
done
grep -E "^FAILED|^ERROR|^E  " $O/${P}_pytest_gpu.log | cut -c1-240 | head -30
show() {
  python - <<PY
import json
try:
    d = json.load(open("$O/${P}_bench_$1.json"))
    e = d.get("reference_cuda_eager") or {}
    print("$1: value %.1f e2e %.1f | single %s / %s | tf32 %s / %s | eager fp32 %s autocast %s | cpu %s | roofline %.3f (tf32 %s) memread %.1f us share %.2f | clocks %s" % (
        d["value"], d["e2e"]["value"], d.get("value_single_session"), d.get("e2e_single_session"), d.get("value_tf32"), d.get("e2e_tf32"),
        e.get("fp32", {}).get("value"), e.get("autocast_fp16", {}).get("value"), d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 3), d["roofline"]["frac"],
        (d.get("roofline_tf32") or {}).get("frac"), d["roofline_memory_read"]["avg_call_us"], d["roofline_memory_read"]["share_of_step"], d["clocks"]))
except Exception as ex:
    print("$1 failed:", ex); print(open("$O/${P}_bench_$1.err").read()[-800:])
PY
}
echo "== 2. default line + reference arm (driver's flags)"
(timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${P}_bench_default.json 2> $O/${P}_bench_default.err); show default
(timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/${P}_bench_reference.json 2> $O/${P}_bench_reference.err); python -c "
import json; d=json.load(open('gpurun_out/r02c14_bench_reference.json')); print('reference arm: value %.2f frames/s, ms_per_step %.0f, cores %s' % (d['value'], d['ms_per_step'], d['cpu_baseline']['cores']))"
echo "== 3. cfg3 / cfg5 lines"
(timeout 500 python bench.py --config cfg3 --steps 3 --warmup 3 --extra-steps 2 > $O/${P}_bench_cfg3.json 2> $O/${P}_bench_cfg3.err); show cfg3
(timeout 800 python bench.py --config cfg5 --steps 2 --warmup 3 --extra-steps 2 > $O/${P}_bench_cfg5.json 2> $O/${P}_bench_cfg5.err); show cfg5
echo "== 4. launch list of a steady-state lock-step region"
(timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${P}_lockstep_launches.csv \
  python tools/prof_lockstep_region.py > $O/${P}_prof_region.log 2>&1); tail -2 $O/${P}_prof_region.log
python tools/ncu_summary.py launches $O/${P}_lockstep_launches.csv > $O/${P}_lockstep_launch_list.txt 2>&1; head -30 $O/${P}_lockstep_launch_list.txt
echo "== done"
