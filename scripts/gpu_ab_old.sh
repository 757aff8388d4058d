#!/bin/bash
# A/B of the LK kernel against the library built from an older commit (lib/libklt_b200_old.so), same box, same run
OUT=gpurun_out/${1:-abold}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/pytest.log
for i in 1 2; do
echo "== new (scheduled)"; timeout 300 python scripts/lk_time.py 2>&1 | grep "B=" | tee -a $OUT/new.log
echo "== new (KLT_LK_NOSCHED=1)"; KLT_LK_NOSCHED=1 timeout 300 python scripts/lk_time.py 2>&1 | grep "B=" | tee -a $OUT/new_nosched.log
echo "== old"; KLT_LIB_PATH=$PWD/visual-odom-pipeline_b200/lib/libklt_b200_old.so timeout 300 python scripts/lk_time.py 2>&1 | grep "B=" | tee -a $OUT/old.log
done
timeout 200 python scripts/lk_timeline.py 2>&1 | grep -v "^  [0-9-]*/" | head -12 | cut -c1-200 | tee $OUT/timeline_win21.log
