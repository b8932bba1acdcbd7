#!/bin/bash
# Round 2, GPU call 6: tap-reuse mode for 3x3 convs (T9), channel-tile-fastest tile order; A/Bs; ncu of the expansion conv.
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/r02c6_pytest_gpu.log
for f in tests/test_gpu_fp16.py tests/test_gpu_ops.py tests/test_gpu_network.py tests/test_gpu_zz_lockstep.py tests/test_gpu_z_s2m.py tests/test_gpu_zzz_fullsize.py; do
  echo "=== $f" >> $O/r02c6_pytest_gpu.log
  (timeout 600 python -m pytest $f -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "^$" | tail -40 >> $O/r02c6_pytest_gpu.log)
  echo "$f: $(grep -E 'passed|failed|error' $O/r02c6_pytest_gpu.log | tail -1)"
done
grep -E "^FAILED|^ERROR|^E  " $O/r02c6_pytest_gpu.log | cut -c1-240 | head -30
echo "== 2. per-layer tables: default / tap reuse off / round-1 tile order"
(timeout 120 python tools/lockstep_layer_times.py 4 fp16 > $O/r02c6_layers_lockstep4.log 2>&1); head -6 $O/r02c6_layers_lockstep4.log; grep "+res\|3x3" $O/r02c6_layers_lockstep4.log | head -40
(MIVOS_CONV_TAP3=0 timeout 120 python tools/lockstep_layer_times.py 4 fp16 > $O/r02c6_layers_lockstep4_notap3.log 2>&1); head -6 $O/r02c6_layers_lockstep4_notap3.log
(MIVOS_CONV_TILE_ORDER=m timeout 120 python tools/lockstep_layer_times.py 4 fp16 > $O/r02c6_layers_lockstep4_orderm.log 2>&1); head -6 $O/r02c6_layers_lockstep4_orderm.log; grep "+res" $O/r02c6_layers_lockstep4_orderm.log | head -6
(timeout 120 python tools/lockstep_layer_times.py 1 fp16 > $O/r02c6_layers_lockstep1.log 2>&1); head -6 $O/r02c6_layers_lockstep1.log
(MIVOS_CONV_TAP3=0 timeout 120 python tools/lockstep_layer_times.py 1 fp16 > $O/r02c6_layers_lockstep1_notap3.log 2>&1); head -6 $O/r02c6_layers_lockstep1_notap3.log
echo "== 3. bench"
for cfg in "2 4" "3 4" "1 1"; do
  set -- $cfg
  (timeout 300 python bench.py --clips-per-gpu $1 --lockstep $2 --steps 3 --warmup 2 --skip-cpu-baseline --skip-extras --skip-cuda-eager > $O/r02c6_bench_c$1_l$2.json 2> $O/r02c6_bench_c$1_l$2.err)
  python - <<PY
import json
try:
    d = json.load(open("$O/r02c6_bench_c$1_l$2.json"))
    print("lanes $1 x clips $2: value %.1f e2e %.1f roofline.frac %.3f memread %.1f us launches %d" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_memory_read"]["avg_call_us"], d["gpu_launches"]))
except Exception as e:
    print("lanes $1 x clips $2: failed", e); print(open("$O/r02c6_bench_c$1_l$2.err").read()[-600:])
PY
done
echo "== 4. ncu --set full: expansion conv at n=4 (both tile orders)"
(timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm" -s 1 -c 1 -o $O/r02c6_expand4 python tools/prof_kernels.py expand4 > $O/r02c6_expand4_ncu.log 2>&1)
(MIVOS_CONV_TILE_ORDER=m timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm" -s 1 -c 1 -o $O/r02c6_expand4_orderm python tools/prof_kernels.py expand4 > $O/r02c6_expand4_orderm_ncu.log 2>&1)
echo "== 5. launch list of a timed lock-step region (graph replays on real data)"
(timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 20000 -c 2500 --csv --log-file $O/r02c6_launches.csv \
  python bench.py --clips-per-gpu 1 --lockstep 4 --steps 1 --warmup 1 --skip-cpu-baseline --skip-extras --skip-cuda-eager --skip-roofline > $O/r02c6_bench_under_ncu.log 2>&1)
python tools/ncu_summary.py launches $O/r02c6_launches.csv > $O/r02c6_launch_list.txt 2>&1; head -40 $O/r02c6_launch_list.txt; tail -5 $O/r02c6_bench_under_ncu.log
echo "== done"
