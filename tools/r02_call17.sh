#!/bin/bash
# Round 2, GPU call 17: the fallback decision moved into the selection stage (two-pass protocol: pass A flags a query
# whose lists hold more in-band candidates than it can stage, exact generator, pass B serves the flagged queries):
# memory-read tests (the huge-norm case is the one that needs the fallback), full-size goldens, cfg3 line, the rest.
set -u
mkdir -p gpurun_out
O=gpurun_out
P=r02c17
: > $O/${P}_pytest_gpu.log
for f in tests/test_gpu_memread.py tests/test_gpu_zzz_fullsize.py tests/test_gpu_zz_batched_ops.py; do
  echo "=== $f" >> $O/${P}_pytest_gpu.log
  (timeout 300 python -m pytest $f -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "^$" | tail -40 >> $O/${P}_pytest_gpu.log)
  echo "$f: $(grep -E 'passed|failed|error' $O/${P}_pytest_gpu.log | tail -1)"
done
grep -E "^FAILED|^ERROR|^E  " $O/${P}_pytest_gpu.log | cut -c1-240 | head -20
(timeout 200 python bench.py --config cfg3 --steps 3 --warmup 3 --skip-extras --skip-cuda-eager --skip-cpu-baseline > $O/${P}_bench_cfg3.json 2> $O/${P}_bench_cfg3.err)
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02c17_bench_cfg3.json"))
    print("cfg3: value %.1f e2e %.1f | memread %.1f us share %.2f" % (d["value"], d["e2e"]["value"], d["roofline_memory_read"]["avg_call_us"], d["roofline_memory_read"]["share_of_step"]))
except Exception as ex:
    print("cfg3 failed:", ex); print(open("gpurun_out/r02c17_bench_cfg3.err").read()[-800:])
PY
for f in tests/test_gpu_zz_lockstep.py tests/test_gpu_network.py tests/test_gpu_clients.py; do
  echo "=== $f" >> $O/${P}_pytest_gpu.log
  (timeout 200 python -m pytest $f -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "^$" | tail -40 >> $O/${P}_pytest_gpu.log)
  echo "$f: $(grep -E 'passed|failed|error' $O/${P}_pytest_gpu.log | tail -1)"
done
echo "== done"
