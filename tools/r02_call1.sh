#!/bin/bash
# Round 2, GPU call 1: baseline of the round-1 tree on this round's box + the measurements VERDICT r01
# lists as missing (lock-step launch list / per-layer table, sanitizers, lock-step A/Bs) + the UMMA
# descriptor probe that decides whether one activation box can serve three 3x3 taps.
#   /usr/local/graft/bin/gpurun --timeout 2100 -- 'bash tools/r02_call1.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02c1_smi.txt 2>&1
echo "== 1. UMMA descriptor probe"; (timeout 60 tools/build/umma_probe > $O/r02c1_umma_probe.log 2>&1); echo "rc=$?"; cat $O/r02c1_umma_probe.log
echo "== 2. GPU suite"; (timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -30 > $O/r02c1_pytest_gpu.log); tail -3 $O/r02c1_pytest_gpu.log
echo "== 3. per-layer tables"
(timeout 120 python tools/lockstep_layer_times.py 4 fp16 > $O/r02c1_layers_lockstep4_fp16.log 2>&1); head -8 $O/r02c1_layers_lockstep4_fp16.log
(timeout 120 python tools/lockstep_layer_times.py 1 fp16 > $O/r02c1_layers_lockstep1_fp16.log 2>&1); head -8 $O/r02c1_layers_lockstep1_fp16.log
(timeout 120 python tools/time_phases.py > $O/r02c1_phase_times_fp16.log 2>&1); cat $O/r02c1_phase_times_fp16.log
echo "== 4. lock-step A/Bs"
for cfg in "2 4 0" "3 4 0" "2 6 0" "1 1 0"; do
  set -- $cfg
  (MIVOS_LOCKSTEP_JOINT_QUERY=$3 timeout 150 python bench.py --clips-per-gpu $1 --lockstep $2 --steps 2 --warmup 2 --skip-cpu-baseline \
    > $O/r02c1_bench_c$1_l$2_j$3.json 2> $O/r02c1_bench_c$1_l$2_j$3.err)
  python - <<PY
import json
try:
    d = json.load(open("$O/r02c1_bench_c$1_l$2_j$3.json"))
    print("lanes $1 x clips $2 joint=$3: value %.1f e2e %.1f roofline.frac %.3f memread %.1f us" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_memory_read"]["avg_call_us"]))
except Exception as e:
    print("lanes $1 x clips $2 joint=$3: failed", e)
PY
done
echo "== 5. ncu launch list of the default bench"
(timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02c1_launches.csv \
  python bench.py --steps 1 --warmup 1 --skip-cpu-baseline > $O/r02c1_bench_under_ncu.log 2>&1)
python tools/ncu_summary.py launches $O/r02c1_launches.csv > $O/r02c1_launch_list.txt 2>&1; head -32 $O/r02c1_launch_list.txt
echo "== 6. ncu --set full: memory-read kernels + the output-bound expansion conv"
(timeout 200 ncu --set full --clock-control none --import-source on -k regex:"memread" -s 8 -c 8 -o $O/r02c1_memread python tools/prof_kernels.py memread > $O/r02c1_memread_ncu.log 2>&1)
(timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm" -s 1 -c 2 -o $O/r02c1_expand python tools/prof_kernels.py expand > $O/r02c1_expand_ncu.log 2>&1)
ls -la $O/*.ncu-rep 2>/dev/null
echo "== 7. sanitizers on the smoke pass"
for tool in memcheck synccheck racecheck; do
  (timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" \
    > $O/r02c1_sanitizer_${tool}_smoke.log 2>&1); echo "$tool smoke rc=$?"; tail -3 $O/r02c1_sanitizer_${tool}_smoke.log
done
echo "== done"
