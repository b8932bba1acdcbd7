#!/bin/bash
# Round 2, GPU call 15: top-k 50 variant of the candidate pass back on the 32-column reads (the pipelined reads
# measured slower there): its tests (memory read, full-size cfg3 / cfg5 goldens) and the cfg3 / cfg5 lines again;
# sanitizers (memcheck, synccheck) on smoke() with this round's kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out
P=r02c15
: > $O/${P}_pytest_gpu.log
for f in tests/test_gpu_memread.py tests/test_gpu_zz_batched_ops.py tests/test_gpu_zzz_fullsize.py; do
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
    print("$1: value %.1f e2e %.1f | roofline %.3f memread %.1f us share %.2f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"],
          d["roofline_memory_read"]["avg_call_us"], d["roofline_memory_read"]["share_of_step"]))
except Exception as ex:
    print("$1 failed:", ex); print(open("$O/${P}_bench_$1.err").read()[-800:])
PY
}
(timeout 300 python bench.py --config cfg3 --steps 3 --warmup 3 --skip-extras --skip-cuda-eager --skip-cpu-baseline > $O/${P}_bench_cfg3.json 2> $O/${P}_bench_cfg3.err); show cfg3
(timeout 400 python bench.py --config cfg5 --steps 2 --warmup 3 --skip-extras --skip-cuda-eager --skip-cpu-baseline > $O/${P}_bench_cfg5.json 2> $O/${P}_bench_cfg5.err); show cfg5
echo "== sanitizers on smoke() with this round's kernels (TMA epilogue, tap reuse, warp-per-query selection, PTX emission)"
for tool in memcheck synccheck; do
  (timeout 150 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/${P}_sanitizer_${tool}_smoke.log 2>&1); echo "$tool smoke rc=$?"; tail -3 $O/${P}_sanitizer_${tool}_smoke.log
done
echo "== done"
