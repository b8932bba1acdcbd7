#!/bin/bash
# Round 2, GPU call 16 (2 GPUs): the default bench line under torchrun exactly as the driver launches it (N = 2).
set -u
mkdir -p gpurun_out
O=gpurun_out
P=r02c16
nvidia-smi --query-gpu=index,name --format=csv,noheader
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 6 --warmup 3 \
  > $O/${P}_bench_n2.json 2> $O/${P}_bench_n2.err); tail -c 300 $O/${P}_bench_n2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02c16_bench_n2.json"))
    print("N=2 default: value %.1f e2e %.1f | single %s | tf32 %s | n_gpus %s | ms_per_step %.1f | launches %s | clocks %s" % (
        d["value"], d["e2e"]["value"], d.get("value_single_session"), d.get("value_tf32"), d["n_gpus"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
except Exception as ex:
    print("N=2 failed:", ex); print(open("gpurun_out/r02c16_bench_n2.err").read()[-1500:])
PY
echo "== done"
