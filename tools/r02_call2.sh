#!/bin/bash
# Round 2, GPU call 2: TMA conv epilogue A/B + the new full-size parity tests + the new bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== 1. GPU suite (TMA epilogue on)"; (timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -s 2>&1 | grep -v "^$" | tail -150 > $O/r02c2_pytest_gpu.log); grep -E "passed|failed|FAILED|\[fullsize\]" $O/r02c2_pytest_gpu.log | tail -60
if grep -q "failed" $O/r02c2_pytest_gpu.log; then
  echo "== 1b. conv / network tests with the register epilogue"
  (MIVOS_CONV_TMA_EPILOGUE=0 timeout 600 python -m pytest tests/test_gpu_fp16.py tests/test_gpu_ops.py tests/test_gpu_network.py -m gpu -q -p no:cacheprovider --tb=line 2>&1 | tail -15 > $O/r02c2_pytest_gpu_regepi.log); tail -5 $O/r02c2_pytest_gpu_regepi.log
fi
echo "== 2. per-layer tables, TMA epilogue on / off"
(timeout 120 python tools/lockstep_layer_times.py 4 fp16 > $O/r02c2_layers_lockstep4_tma.log 2>&1); head -6 $O/r02c2_layers_lockstep4_tma.log; grep "+res" $O/r02c2_layers_lockstep4_tma.log | head -8
(MIVOS_CONV_TMA_EPILOGUE=0 timeout 120 python tools/lockstep_layer_times.py 4 fp16 > $O/r02c2_layers_lockstep4_reg.log 2>&1); head -6 $O/r02c2_layers_lockstep4_reg.log
(timeout 120 python tools/lockstep_layer_times.py 1 fp16 > $O/r02c2_layers_lockstep1_tma.log 2>&1); head -6 $O/r02c2_layers_lockstep1_tma.log
echo "== 3. bench (new line)"
(timeout 600 python bench.py --steps 3 --warmup 3 > $O/r02c2_bench.json 2> $O/r02c2_bench.err); tail -c 600 $O/r02c2_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02c2_bench.json"))
    print("value %.1f e2e %.1f | single %s | tf32 %s | eager %s | cpu %s | roofline %.3f memread %.1f us" % (
        d["value"], d["e2e"]["value"], d.get("value_single_session"), d.get("value_tf32"),
        {k: (round(v["value"], 1) if isinstance(v, dict) else v) for k, v in (d.get("reference_cuda_eager") or {}).items()},
        d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 2), d["roofline"]["frac"], d["roofline_memory_read"]["avg_call_us"]))
except Exception as e:
    print("bench line failed:", e)
PY
(timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02c2_bench_reference.json 2> $O/r02c2_bench_reference.err); tail -c 700 $O/r02c2_bench_reference.json; tail -c 300 $O/r02c2_bench_reference.err
echo "== 4. ncu --set full of the expansion conv with the TMA epilogue"
(timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm" -s 1 -c 2 -o $O/r02c2_expand_tma python tools/prof_kernels.py expand > $O/r02c2_expand_ncu.log 2>&1)
echo "== done"
