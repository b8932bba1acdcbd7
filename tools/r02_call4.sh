#!/bin/bash
# Round 2, GPU call 4: everything since call 2 — warp-per-query select + query sets, batched lock-step ops, s2d stems,
# 8-warp TMA epilogue, overlapped forward/backward passes.  One pytest process per file (a sticky CUDA error in one
# file must not take the others down).
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/r02c4_pytest_gpu.log
for f in tests/test_gpu_*.py; do
  echo "=== $f" >> $O/r02c4_pytest_gpu.log
  (timeout 600 python -m pytest $f -m gpu -q -p no:cacheprovider --tb=short -s 2>&1 | grep -v "^$" | tail -60 >> $O/r02c4_pytest_gpu.log)
  echo "$f: $(grep -E 'passed|failed|error' $O/r02c4_pytest_gpu.log | tail -1)"
done
grep -E "^FAILED|^ERROR|\[fullsize\] (cfg3|cfg4|cfg5|2x4|cfg2_single)" $O/r02c4_pytest_gpu.log | cut -c1-260 | tail -50
echo "== 2. per-layer tables"
(timeout 120 python tools/lockstep_layer_times.py 4 fp16 > $O/r02c4_layers_lockstep4.log 2>&1); head -14 $O/r02c4_layers_lockstep4.log; grep "+res\|stem\|upsample" $O/r02c4_layers_lockstep4.log | head -12
(MIVOS_CONV_EPI8=0 timeout 120 python tools/lockstep_layer_times.py 4 fp16 > $O/r02c4_layers_lockstep4_epi4.log 2>&1); head -6 $O/r02c4_layers_lockstep4_epi4.log; grep "+res" $O/r02c4_layers_lockstep4_epi4.log | head -6
(timeout 120 python tools/lockstep_layer_times.py 1 fp16 > $O/r02c4_layers_lockstep1.log 2>&1); head -8 $O/r02c4_layers_lockstep1.log
(timeout 120 python tools/time_phases.py > $O/r02c4_phase_times.log 2>&1); tail -7 $O/r02c4_phase_times.log
echo "== 3. bench"
for cfg in "2 4" "3 4" "1 1"; do
  set -- $cfg
  (timeout 300 python bench.py --clips-per-gpu $1 --lockstep $2 --steps 3 --warmup 2 --skip-cpu-baseline --skip-extras --skip-cuda-eager > $O/r02c4_bench_c$1_l$2.json 2> $O/r02c4_bench_c$1_l$2.err)
  python - <<PY
import json
try:
    d = json.load(open("$O/r02c4_bench_c$1_l$2.json"))
    print("lanes $1 x clips $2: value %.1f e2e %.1f roofline.frac %.3f memread %.1f us launches %d" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_memory_read"]["avg_call_us"], d["gpu_launches"]))
except Exception as e:
    print("lanes $1 x clips $2: failed", e); print(open("$O/r02c4_bench_c$1_l$2.err").read()[-600:])
PY
done
echo "== 4. ncu of the memory read (one lock-step step of 4 clips)"
(timeout 240 ncu --set full --clock-control none --import-source on -k regex:"memread" -s 10 -c 8 -o $O/r02c4_memread python tools/lockstep_layer_times.py 4 fp16 > $O/r02c4_memread_ncu.log 2>&1)
echo "== done"
