#!/bin/bash
# iteration check of the leader/follower LK kernel: timings first (fast feedback), then the GPU tests, then the timeline
OUT=gpurun_out/${1:-r02_iter2}
mkdir -p $OUT
run() { echo "== $*" >> $OUT/lk_time.log; env "$@" timeout 300 python scripts/lk_time.py >> $OUT/lk_time.log 2>&1; }
run KLT_X=0
run KLT_LK_SLOTS=2
run KLT_LK_SLOTS=1
cat $OUT/lk_time.log
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
KLT_LIB_PATH=$PWD/visual-odom-pipeline_b200/lib/libklt_b200_tl.so timeout 300 python scripts/lk_timeline.py > $OUT/timeline.log 2>&1
cat $OUT/timeline.log
