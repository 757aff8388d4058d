#!/bin/bash
OUT=gpurun_out/${1:-e2e}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/pytest.log
run() { python bench.py --steps 400 --no-cpu-baseline --workload $2 > $OUT/b.json 2>$OUT/bench.err; python -c "
import json; d=json.load(open('$OUT/b.json')); print('$1 $2: ms/step %.4f'%d['ms_per_step'], 'e2e ms %.4f'%d['e2e']['ms_per_step'], 'e2e kp/s %.2f M'%(d['e2e']['value']/1e6))"; }
for wl in kitti kitti_ref_params; do
run order $wl
KLT_LK_NOORDER=1 run noorder $wl
run order $wl
KLT_LK_NOORDER=1 run noorder $wl
done
