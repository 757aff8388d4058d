#!/bin/bash
TAG=${1:-pp}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in ${VARIANTS:-1}; do
KLT_PYR_RING=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:pyr_down -s 2 -c 1 -o $OUT/pyr_malaga_v$v -f python scripts/prof_pyr.py 768 1024 256 > $OUT/prof_malaga_v$v.log 2>&1
KLT_PYR_RING=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:pyr_down -s 2 -c 1 -o $OUT/pyr_kitti_v$v -f python scripts/prof_pyr.py 376 1241 310 > $OUT/prof_kitti_v$v.log 2>&1
done
ls -la $OUT
