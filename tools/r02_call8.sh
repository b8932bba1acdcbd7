#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
(timeout 300 python tools/tile_sweep.py > $O/r02c8_tile_sweep_fp16.log 2>&1); cat $O/r02c8_tile_sweep_fp16.log
echo "== 4 epilogue warps on the 128-wide tiles"
(MIVOS_CONV_EPI8=0 timeout 200 python tools/tile_sweep.py expand > $O/r02c8_tile_sweep_expand_epi4.log 2>&1); cat $O/r02c8_tile_sweep_expand_epi4.log
echo "== register epilogue"
(MIVOS_CONV_TMA_EPILOGUE=0 timeout 200 python tools/tile_sweep.py expand > $O/r02c8_tile_sweep_expand_regepi.log 2>&1); cat $O/r02c8_tile_sweep_expand_regepi.log
