#!/bin/bash
# One short GPU call (round-1 re-entry, 7 GPU-minutes left): full GPU suite (new S2M / egress /
# lock-step tests included, no -x so every failure is listed), S2M timing, one lock-step bench line.
mkdir -p gpurun_out
(timeout 170 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -s 2>&1 | tail -220 > gpurun_out/r01b_pytest_gpu.log)
tail -4 gpurun_out/r01b_pytest_gpu.log
(timeout 45 python tools/s2m_time.py fp16 > gpurun_out/r01b_s2m_time.log 2>&1); cat gpurun_out/r01b_s2m_time.log
(timeout 110 python bench.py --clips-per-gpu 1 --lockstep 4 --steps 2 --warmup 2 --skip-cpu-baseline > gpurun_out/r01b_bench_lockstep4.json 2> gpurun_out/r01b_bench_lockstep4.err); tail -c 1500 gpurun_out/r01b_bench_lockstep4.json; tail -3 gpurun_out/r01b_bench_lockstep4.err
