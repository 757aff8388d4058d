#!/bin/bash
# Timeline / phase diagnostics of the single-pair LK launch (needs lib/libklt_b200_tl.so: build.py --variant tl --extra -DKLT_LK_TIMELINE)
OUT=gpurun_out/${1:-r02_lk_tl}
mkdir -p $OUT
TL=$PWD/visual-odom-pipeline_b200/lib/libklt_b200_tl.so
for B in -1 7; do
  echo "== KLT_LK_BUDGET=$B" >> $OUT/timeline.log
  KLT_LIB_PATH=$TL KLT_LK_BUDGET=$B timeout 300 python scripts/lk_timeline.py >> $OUT/timeline.log 2>&1
done
echo "== phases BUDGET=7" >> $OUT/timeline.log
KLT_LIB_PATH=$TL KLT_LK_BUDGET=7 timeout 300 python scripts/lk_phases.py >> $OUT/timeline.log 2>&1
cat $OUT/timeline.log
