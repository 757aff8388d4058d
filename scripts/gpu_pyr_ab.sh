#!/bin/bash
# One-launch pyramid build vs one launch per level: parity tests, then timings on inputs larger than L2.
OUT=gpurun_out/${1:-r02_pyr_ab}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "pyramid or pipeline or smoke or golden" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
echo "== KLT_PYR_ONE_LAUNCH=0" > $OUT/pyr_time.log; KLT_PYR_ONE_LAUNCH=0 timeout 300 python scripts/pyr_time.py >> $OUT/pyr_time.log 2>&1
echo "== one launch (default)" >> $OUT/pyr_time.log; timeout 300 python scripts/pyr_time.py >> $OUT/pyr_time.log 2>&1
cat $OUT/pyr_time.log
