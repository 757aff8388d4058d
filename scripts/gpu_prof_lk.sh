#!/bin/bash
# ncu --set full of the LK kernels: single KITTI pair (and optionally the batched shape).  usage: gpu_prof_lk.sh <tag> [regex] [mode]
OUT=gpurun_out/${1:-r02_prof_lk}
RX=${2:-lk_}
MODE=${3:-lk}
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RX -s 2 -c 2 -o $OUT/prof_$MODE -f \
    python scripts/prof_target.py $MODE > $OUT/prof_$MODE.log 2>&1
tail -3 $OUT/prof_$MODE.log
ls -la $OUT
