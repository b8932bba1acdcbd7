#!/bin/bash
# Round 2, GPU call 11 (call 10 lost its box before anything came back; its steps are re-issued as short calls):
# GPU suite, memory-read A/B (emission in PTX / barrier back-off), joint query pass A/B, per-layer table.
set -u
mkdir -p gpurun_out
O=gpurun_out
P=r02c11
echo "box: $(nproc) cpus, $(free -g | awk '/Mem:/{print $2" GB RAM, "$7" GB available"}'), $(nvidia-smi --query-gpu=name,memory.total --format=csv,noheader)"
echo "== 1. GPU suite"
: > $O/${P}_pytest_gpu.log
for f in tests/test_gpu_*.py; do
  echo "=== $f" >> $O/${P}_pytest_gpu.log
  (timeout 400 python -m pytest $f -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "^$" | tail -40 >> $O/${P}_pytest_gpu.log)
  echo "$f: $(grep -E 'passed|failed|error' $O/${P}_pytest_gpu.log | tail -1)"
done
grep -E "^FAILED|^ERROR|^E  " $O/${P}_pytest_gpu.log | cut -c1-240 | head -30
echo "== 2. A/B (2 lanes x 4 clips, steps 3)"
run_ab() {  # name, env...
  local name=$1; shift
  (env "$@" timeout 200 python bench.py --clips-per-gpu 2 --lockstep 4 --steps 3 --warmup 2 --skip-cpu-baseline --skip-extras --skip-cuda-eager > $O/${P}_ab_$name.json 2> $O/${P}_ab_$name.err)
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
echo "== done"
