#!/bin/bash
OUT=gpurun_out/${1:-ab8}; mkdir -p $OUT
timeout 200 python -m pytest tests -m gpu -x -q -k "pyramid or native" 2>&1 | tail -2
echo "== 4 CTAs/SM"; KLT_PYR_RING=4 timeout 200 python scripts/pyr_time.py 2>&1 | grep -v "B=2:" | tee $OUT/v4.log
for r in 8 12 16 24 32; do echo "== rows $r"; KLT_PYR_ROWS=$r timeout 200 python scripts/pyr_time.py 2>&1 | grep -v "B=2:" | tee $OUT/rows$r.log; done
