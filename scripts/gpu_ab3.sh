#!/bin/bash
TAG=${1:-ab3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q -k "pyramid or native" > $OUT/pytest_pyr.log 2>&1; tail -3 $OUT/pytest_pyr.log
for v in ${VARIANTS:-1}; do
KLT_PYR_RING=$v timeout 300 python scripts/pyr_time.py > $OUT/pyr_time_v$v.log 2>&1; echo "variant $v"; cat $OUT/pyr_time_v$v.log
done
