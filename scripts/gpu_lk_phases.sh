#!/bin/bash
# Phase diagnostics of the single-pair LK launch (builds lib/libklt_b200_phases.so here if missing: -DKLT_LK_PHASES)
OUT=gpurun_out/${1:-r02_lk_phases}
mkdir -p $OUT
LIB=$PWD/visual-odom-pipeline_b200/lib/libklt_b200_phases.so
[ -f $LIB ] || python visual-odom-pipeline_b200/build.py --variant phases --extra -DKLT_LK_PHASES
KLT_LIB_PATH=$LIB timeout 300 python scripts/lk_phases.py "${@:2}" > $OUT/phases.log 2>&1
cat $OUT/phases.log
