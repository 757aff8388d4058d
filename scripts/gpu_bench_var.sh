#!/bin/bash
# bench.py A/B of library variants in lib/ on one box: args = tag, workloads (comma separated), then variant suffixes ("" = product build)
OUT=gpurun_out/${1:-benchvar}; mkdir -p $OUT; shift
WLS=${1//,/ }; shift
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/pytest.log
for rep in 1 2; do
for wl in $WLS; do
for v in "$@"; do
  KLT_LIB_PATH=$PWD/visual-odom-pipeline_b200/lib/libklt_b200${v}.so python bench.py --steps 300 --no-cpu-baseline --workload $wl > $OUT/b_${wl}${v}.json 2>$OUT/bench.err; python -c "
import json; d=json.load(open('$OUT/b_${wl}${v}.json')); print('$wl \'$v\': ms/step %.4f'%d['ms_per_step'], 'lk %.4f'%d['kernel_ms']['lk'], 'e2e ms %.4f'%d['e2e']['ms_per_step'], 'pipelined %.2f'%(d['pipelined']['keypoints_per_sec']/1e6), 'batched %.2f M/s'%(d['batched_lk']['keypoints_per_sec']/1e6), d['parity']['bit_exact'])"
done; done; done
