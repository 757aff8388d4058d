#!/bin/bash
# ncu --set full of the LK kernel: single KITTI pair (mode lk) or the batched shape (mode batch).  usage: gpu_prof_lk.sh <tag> [mode] [env...]
OUT=gpurun_out/${1:-r02_prof_lk}
MODE=${2:-lk}
shift; shift
mkdir -p $OUT
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:lk_ -s 2 -c 1 -o $OUT/prof_$MODE -f \
    python scripts/prof_target.py $MODE > $OUT/prof_$MODE.log 2>&1
tail -3 $OUT/prof_$MODE.log
ls -la $OUT
