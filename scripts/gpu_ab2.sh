#!/bin/bash
TAG=${1:-ab2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
python scripts/pyr_time.py > $OUT/pyr_time.log 2>&1; cat $OUT/pyr_time.log
KLT_TRACE=1 python scripts/e2e_trace.py > $OUT/e2e_trace.log 2>&1; grep -E "median" -B2 $OUT/e2e_trace.log
python scripts/lk_time.py > $OUT/lk_time.log 2>&1; cat $OUT/lk_time.log
