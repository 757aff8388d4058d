#!/bin/bash
# A/B of library variants in lib/ (same box, same run): args = tag, then variant suffixes ("" = product build)
OUT=gpurun_out/${1:-abvar}; mkdir -p $OUT; shift
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/pytest.log
for i in 1 2; do
for v in "$@"; do
  lib=$PWD/visual-odom-pipeline_b200/lib/libklt_b200${v}.so
  echo "== variant '${v}'"; KLT_LIB_PATH=$lib timeout 300 python scripts/lk_time.py 2>&1 | grep "B=" | tee -a $OUT/var${v}.log
done
done
