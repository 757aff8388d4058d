#!/bin/bash
# A/B of two builds of the library on one box (KLT_LIB_PATH): LK timings and the bench line.  usage: gpu_ab_lib.sh <tag> <variant.so>
OUT=gpurun_out/${1:-r02_ab}
VAR=$PWD/visual-odom-pipeline_b200/lib/${2:-libklt_b200_minb6.so}
mkdir -p $OUT
for rep in 1 2; do
  echo "== default (rep $rep)" >> $OUT/lk_time.log; timeout 300 python scripts/lk_time.py >> $OUT/lk_time.log 2>&1
  echo "== $2 (rep $rep)" >> $OUT/lk_time.log; KLT_LIB_PATH=$VAR timeout 300 python scripts/lk_time.py >> $OUT/lk_time.log 2>&1
done
cat $OUT/lk_time.log
python bench.py --no-detection --no-cpu-baseline --no-sharded-batch > $OUT/bench_default.json 2> $OUT/bench.err
KLT_LIB_PATH=$VAR python bench.py --no-detection --no-cpu-baseline --no-sharded-batch > $OUT/bench_variant.json 2>> $OUT/bench.err
python - $OUT <<'PY'
import json, sys
for f in ("bench_default.json", "bench_variant.json"):
    try:
        d = json.loads(open("%s/%s" % (sys.argv[1], f)).read().strip().splitlines()[-1])
        print(f, "value %.2f M ms/step %.4f e2e %.4f ms pageable %.4f ms kernel_ms %s pipelined %.4f batched %.2f M"
              % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e_pageable"]["ms_per_step"], d["kernel_ms"],
                 d["pipelined"]["ms_per_step"], d["batched_lk"]["keypoints_per_sec"] / 1e6))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $OUT/bench.err
