#!/bin/bash
# Round 2, GPU call 13: memory read at top-k 50 (cfg3 / cfg5): no final cap (the exact fallback ran for 3/4 of the
# 32-query tiles at cfg3), pair threshold between the column halves, selection stage pre-filters against the final
# shared threshold.  Tests of the read, the networks and the full-size goldens; bench lines of cfg3, cfg5, cfg2 (quick).
set -u
mkdir -p gpurun_out
O=gpurun_out
P=r02c13
: > $O/${P}_pytest_gpu.log
for f in tests/test_gpu_memread.py tests/test_gpu_zz_batched_ops.py tests/test_gpu_network.py tests/test_gpu_zz_lockstep.py tests/test_gpu_zzz_fullsize.py; do
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
    print("$1: value %.1f e2e %.1f | single %s / %s | tf32 %s / %s | eager fp32 %s autocast %s | cpu %s | roofline %.3f memread %.1f us share %.2f" % (
        d["value"], d["e2e"]["value"], d.get("value_single_session"), d.get("e2e_single_session"), d.get("value_tf32"), d.get("e2e_tf32"),
        e.get("fp32", {}).get("value"), e.get("autocast_fp16", {}).get("value"), d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 3), d["roofline"]["frac"],
        d["roofline_memory_read"]["avg_call_us"], d["roofline_memory_read"]["share_of_step"]))
except Exception as ex:
    print("$1 failed:", ex); print(open("$O/${P}_bench_$1.err").read()[-800:])
PY
}
(timeout 300 python bench.py --clips-per-gpu 2 --lockstep 4 --steps 3 --warmup 2 --skip-cpu-baseline --skip-extras --skip-cuda-eager > $O/${P}_bench_cfg2_quick.json 2> $O/${P}_bench_cfg2_quick.err); show cfg2_quick
(timeout 600 python bench.py --config cfg3 --steps 3 --warmup 3 --extra-steps 2 > $O/${P}_bench_cfg3.json 2> $O/${P}_bench_cfg3.err); show cfg3
(timeout 900 python bench.py --config cfg5 --steps 2 --warmup 3 --extra-steps 2 > $O/${P}_bench_cfg5.json 2> $O/${P}_bench_cfg5.err); show cfg5
echo "== done"
