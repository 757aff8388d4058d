#!/bin/bash
# iteration check: GPU tests, pageable vs pinned e2e (stager on/off), LK WPP x BUDGET combos
OUT=gpurun_out/${1:-r02_iter1}
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
echo "== e2e stager on" > $OUT/e2e.log; python scripts/e2e_trace.py 2>> $OUT/e2e.log
echo "== e2e stager off (driver stages pageable memory)" >> $OUT/e2e.log; KLT_NO_STAGER=1 python scripts/e2e_trace.py 2>> $OUT/e2e.log
cat $OUT/e2e.log
run() { echo "== $*" >> $OUT/lk_time.log; env "$@" timeout 300 python scripts/lk_time.py 2>&1 | grep "B=1 " >> $OUT/lk_time.log; }
run KLT_LK_BUDGET=-1
run KLT_LK_BUDGET=7
run KLT_LK_BUDGET=7 KLT_LK_WPP=2
run KLT_LK_BUDGET=5 KLT_LK_WPP=2
run KLT_LK_BUDGET=4 KLT_LK_WPP=2 KLT_LK_RESUME_BLOCKS=148
run KLT_LK_BUDGET=5
cat $OUT/lk_time.log
