#!/bin/bash
# First GPU call of the next session (everything here was written after this round's GPU budget was
# spent, or could not be profiled within it).  Run under gpurun from the repo root; ~25 minutes with the
# sanitizer passes (section 4), ~8 without.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/next_gpu_session.sh'
set -u
mkdir -p gpurun_out
# 1. the GPU suite including the A/B options that only ran on the CPU emulator so far
MIVOS_UNVALIDATED=1 timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -40 > gpurun_out/next_pytest_gpu.log
tail -3 gpurun_out/next_pytest_gpu.log
# 2. lock-step A/Bs on cfg-2: joint query pass, lane / clip counts around the default (2 x 4)
for cfg in "2 4 0" "2 4 1" "3 4 0" "3 3 0" "2 6 0"; do
  set -- $cfg
  MIVOS_LOCKSTEP_JOINT_QUERY=$3 timeout 100 python bench.py --clips-per-gpu $1 --lockstep $2 --steps 2 --warmup 2 --skip-cpu-baseline \
    > gpurun_out/next_bench_c$1_l$2_j$3.json 2> gpurun_out/next_bench_c$1_l$2_j$3.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/next_bench_c$1_l$2_j$3.json"))
    print("lanes $1 x clips $2 joint=$3: value %.1f e2e %.1f roofline.frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
except Exception as e:
    print("lanes $1 x clips $2 joint=$3: failed", e)
PY
done
# 3. launch list of the default bench (kernel shares of the lock-step step) and one full capture of the
#    kernels added in the second session of round 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/next_launches.csv \
  python bench.py --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/next_bench_under_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/next_launches.csv > gpurun_out/next_launch_list.txt 2>&1; head -30 gpurun_out/next_launch_list.txt
ncu --set full --clock-control none --import-source on -k regex:"gather_dilated|avgpool|upsample_bilinear|overlay|upsample_to_plane" -c 12 \
  -o gpurun_out/next_s2m_kernels python tools/s2m_time.py fp16 > gpurun_out/next_s2m_under_ncu.log 2>&1
# 4. sanitizers (none were run in round 1): memcheck + racecheck + synccheck on the smoke
#    pass (small shapes: every kernel of the propagation path once) and on the S2M / egress operator tests
for tool in memcheck racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" \
    > gpurun_out/next_sanitizer_${tool}_smoke.log 2>&1; echo "$tool smoke rc=$?"; tail -2 gpurun_out/next_sanitizer_${tool}_smoke.log
done
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_z_s2m.py tests/test_gpu_y_egress.py -m gpu -q -x \
  -k "gather or avgpool or upsample or overlay or stem" > gpurun_out/next_sanitizer_memcheck_ops.log 2>&1; echo "memcheck ops rc=$?"
