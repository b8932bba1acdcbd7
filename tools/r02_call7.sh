#!/bin/bash
# Round 2, GPU call 7: epilogue ILP + 8-warp routing; real-data memory-read kernel times; ncu of the big 3x3 layer.
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/r02c7_pytest_gpu.log
for f in tests/test_gpu_fp16.py tests/test_gpu_ops.py tests/test_gpu_network.py tests/test_gpu_zz_lockstep.py tests/test_gpu_z_s2m.py tests/test_gpu_zzz_fullsize.py; do
  echo "=== $f" >> $O/r02c7_pytest_gpu.log
  (timeout 600 python -m pytest $f -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "^$" | tail -40 >> $O/r02c7_pytest_gpu.log)
  echo "$f: $(grep -E 'passed|failed|error' $O/r02c7_pytest_gpu.log | tail -1)"
done
grep -E "^FAILED|^ERROR|^E  " $O/r02c7_pytest_gpu.log | cut -c1-240 | head -30
echo "== 2. per-layer tables"
(timeout 120 python tools/lockstep_layer_times.py 4 fp16 > $O/r02c7_layers_lockstep4.log 2>&1); head -6 $O/r02c7_layers_lockstep4.log; grep "+res" $O/r02c7_layers_lockstep4.log | head -8
(timeout 120 python tools/lockstep_layer_times.py 1 fp16 > $O/r02c7_layers_lockstep1.log 2>&1); head -6 $O/r02c7_layers_lockstep1.log; grep "+res" $O/r02c7_layers_lockstep1.log | head -4
(timeout 120 python tools/lockstep_layer_times.py 4 tf32 > $O/r02c7_layers_lockstep4_tf32.log 2>&1); head -6 $O/r02c7_layers_lockstep4_tf32.log
echo "== 3. bench"
for cfg in "2 4" "3 4" "1 1"; do
  set -- $cfg
  (timeout 300 python bench.py --clips-per-gpu $1 --lockstep $2 --steps 3 --warmup 2 --skip-cpu-baseline --skip-extras --skip-cuda-eager > $O/r02c7_bench_c$1_l$2.json 2> $O/r02c7_bench_c$1_l$2.err)
  python - <<PY
import json
try:
    d = json.load(open("$O/r02c7_bench_c$1_l$2.json"))
    print("lanes $1 x clips $2: value %.1f e2e %.1f roofline.frac %.3f memread %.1f us launches %d" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_memory_read"]["avg_call_us"], d["gpu_launches"]))
except Exception as e:
    print("lanes $1 x clips $2: failed", e); print(open("$O/r02c7_bench_c$1_l$2.err").read()[-600:])
PY
done
echo "== 4. real-data kernel times of the memory read: eager lock-step run under ncu (launch list of the memread kernels)"
(MIVOS_GRAPH=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"memread" -s 600 -c 120 --csv --log-file $O/r02c7_memread_launches.csv \
  python bench.py --clips-per-gpu 1 --lockstep 4 --steps 1 --warmup 1 --skip-cpu-baseline --skip-extras --skip-cuda-eager --skip-roofline > $O/r02c7_bench_under_ncu.log 2>&1)
python tools/ncu_summary.py launches $O/r02c7_memread_launches.csv 2>&1 | head -12
echo "== 5. ncu --set full: 3x3 256->256 @120x216 n=1 with tap reuse (dominant kernel: traffic + tensor pipe)"
(timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm" -s 0 -c 3 -o $O/r02c7_conv3x3 python tools/prof_kernels.py conv > $O/r02c7_conv3x3_ncu.log 2>&1)
(timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm" -s 1 -c 1 -o $O/r02c7_expand4 python tools/prof_kernels.py expand4 > $O/r02c7_expand4_ncu.log 2>&1)
echo "== done"
