#!/bin/bash
# A/B of the LK hand-off on one box: parity tests, single-pair / batched timings, bench line.
OUT=gpurun_out/${1:-r02_lk_ab}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
run() { echo "== $*" >> $OUT/lk_time.log; env "$@" timeout 300 python scripts/lk_time.py >> $OUT/lk_time.log 2>&1; }
run KLT_LK_BUDGET=-1
run KLT_LK_BUDGET=7
cat $OUT/lk_time.log
KLT_LK_BUDGET=-1 python bench.py --no-detection --no-cpu-baseline > $OUT/bench_nohandoff.json 2> $OUT/bench.err
python bench.py --no-detection > $OUT/bench.json 2>> $OUT/bench.err
python - <<'PY'
import json,sys
for f in ("bench_nohandoff.json","bench.json"):
    try:
        d=json.loads(open("gpurun_out/%s/%s" % (sys.argv[1] if len(sys.argv)>1 else "r02_lk_ab6", f)).read().strip().splitlines()[-1])
        print(f, "value %.2f M ms/step %.4f e2e %.2f M (%.4f ms) kernel_ms %s pipelined %.2f M batched %.2f M parity %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["kernel_ms"], d["pipelined"]["keypoints_per_sec"]/1e6, d["batched_lk"]["keypoints_per_sec"]/1e6, d["parity"]))
    except Exception as e: print(f, "ERR", e)
PY
tail -5 $OUT/bench.err
