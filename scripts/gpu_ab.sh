#!/bin/bash
# A/B timing of the LK kernel for each warps-per-point setting + host-path trace.  bash scripts/gpu_ab.sh <tag>
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for w in 1 2 4; do
  KLT_LK_WPP=$w python scripts/lk_time.py > $OUT/lk_time_wpp$w.log 2>&1
  KLT_LK_WPP=$w python scripts/lk_cycles.py > $OUT/lk_cycles_wpp$w.log 2>&1
done
python scripts/lk_time.py > $OUT/lk_time_auto.log 2>&1
KLT_TRACE=1 python scripts/e2e_trace.py > $OUT/e2e_trace.log 2>&1
tail -n 8 $OUT/e2e_trace.log
cat $OUT/lk_time_wpp*.log $OUT/lk_cycles_wpp*.log
