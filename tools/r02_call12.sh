#!/bin/bash
# Round 2, GPU call 12: bench lines of cfg3 / cfg4 (ours + reference arm), launch list of a steady-state lock-step
# region, ncu --set full of the memory-read kernels on real data and of the kernels outside the top list.
set -u
mkdir -p gpurun_out
O=gpurun_out
P=r02c12
echo "== 1. bench lines of cfg3 / cfg4 (ours + reference arm)"
for c in cfg3 cfg4; do
  (timeout 600 python bench.py --config $c --steps 3 --warmup 3 --extra-steps 2 > $O/${P}_bench_$c.json 2> $O/${P}_bench_$c.err)
  python - <<PY
import json
try:
    d = json.load(open("$O/${P}_bench_$c.json"))
    e = d.get("reference_cuda_eager") or {}
    print("$c: value %.1f e2e %.1f | single %s / %s | tf32 %s / %s | eager fp32 %s | cpu %s | roofline %.3f memread %.1f us share %.2f" % (
        d["value"], d["e2e"]["value"], d.get("value_single_session"), d.get("e2e_single_session"), d.get("value_tf32"), d.get("e2e_tf32"),
        e.get("fp32", {}).get("value"), d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 3), d["roofline"]["frac"],
        d["roofline_memory_read"]["avg_call_us"], d["roofline_memory_read"]["share_of_step"]))
except Exception as ex:
    print("$c failed:", ex); print(open("$O/${P}_bench_$c.err").read()[-800:])
PY
  (timeout 400 python bench.py --impl reference --config $c --steps 1 --warmup 1 > $O/${P}_bench_${c}_reference.json 2> $O/${P}_bench_${c}_reference.err); tail -c 260 $O/${P}_bench_${c}_reference.json; echo
done
echo "== 2. launch list of a steady-state lock-step region (eager launches, all kernels)"
(MIVOS_GRAPH=0 timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2600 --csv --log-file $O/${P}_lockstep_launches.csv \
  python bench.py --clips-per-gpu 1 --lockstep 4 --steps 1 --warmup 1 --skip-cpu-baseline --skip-extras --skip-cuda-eager --skip-roofline > $O/${P}_bench_under_ncu2.log 2>&1)
python tools/ncu_summary.py launches $O/${P}_lockstep_launches.csv > $O/${P}_lockstep_launch_list.txt 2>&1; head -24 $O/${P}_lockstep_launch_list.txt
echo "== 3. ncu --set full: memory-read kernels of a 4-clip lock-step call late in the clip (real data)"
(MIVOS_GRAPH=0 timeout 420 ncu --set full --clock-control none --import-source on -k regex:"memread_(tc|select)" -s 160 -c 4 -o $O/${P}_memread_real \
  python bench.py --clips-per-gpu 1 --lockstep 4 --steps 1 --warmup 1 --skip-cpu-baseline --skip-extras --skip-cuda-eager --skip-roofline > $O/${P}_bench_under_ncu3.log 2>&1)
ls -la $O/${P}_memread_real.ncu-rep
echo "== 4. ncu --set full: S2M kernels, overlay, split-K epilogue"
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gather_dilated|halo_avgpool|upsample_bilinear|overlay|splitk_epilogue" -c 8 -o $O/${P}_misc \
  python tools/prof_misc.py > $O/${P}_misc_ncu.log 2>&1); tail -2 $O/${P}_misc_ncu.log; ls -la $O/${P}_misc.ncu-rep
echo "== done"
