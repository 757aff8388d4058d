#!/bin/bash
OUT=gpurun_out/${1:-e2e}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/pytest.log
run() { python bench.py --steps 300 --no-cpu-baseline --workload kitti > $OUT/b.json 2>$OUT/bench.err; python -c "
import json; d=json.load(open('$OUT/b.json')); print('$1: ms/step %.4f'%d['ms_per_step'], 'e2e ms %.4f'%d['e2e']['ms_per_step'], 'e2e kp/s %.2f M'%(d['e2e']['value']/1e6))"; }
run two_streams+direct
KLT_ONE_COPY_STREAM=1 run one_stream+direct
KLT_ONE_COPY_STREAM=1 KLT_NO_DIRECT_OUT=1 run neither
run two_streams+direct
python scripts/e2e_trace.py 2>&1 | tail -3
