#!/bin/bash
# per-point latency model (debug flag 0x100) for library variants: args = tag, then variant suffixes
OUT=gpurun_out/${1:-cyc}; mkdir -p $OUT; shift
for v in "$@"; do
  for w in win21 win31; do
    echo "== variant '${v}' $w"; KLT_LIB_PATH=$PWD/visual-odom-pipeline_b200/lib/libklt_b200${v}.so timeout 300 python scripts/lk_cycles.py $w 2>&1 | tee -a $OUT/cyc${v}.log
  done
done
