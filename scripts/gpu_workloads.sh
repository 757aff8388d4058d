#!/bin/bash
# bench.py over the other BASELINE configs (parity-test shapes; reported in DESIGN.md, not bench lines)
OUT=gpurun_out/${1:-workloads}; mkdir -p $OUT
for wl in parking malaga stress4k; do
  steps=300; [ $wl = stress4k ] && steps=20
  timeout 900 python bench.py --workload $wl --steps $steps --no-detection --no-sharded-batch > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  python -c "
import json; d=json.load(open('$OUT/bench_$wl.json')); print('$wl: value %.2fM kp/s, ms/step %.4f'%(d['value']/1e6,d['ms_per_step']), 'kernel_ms', {k:round(v,4) for k,v in d['kernel_ms'].items()}, 'e2e %.4f ms'%d['e2e']['ms_per_step'], 'cpu %.3fM (%d thr) %.3f ms, 1thr %.2f ms'%(d['cpu_baseline']['value']/1e6,d['cpu_baseline']['cores'],d['cpu_baseline']['ms_per_call_median'],d['cpu_baseline']['ms_per_call_single_thread']), 'parity', d['parity'], 'pyr', round(d['roofline']['frac'],3), round(d['roofline']['whole_pyramid']['frac'],3), 'iters', round(d['lk_roofline']['iters_per_point'],2), 'pageable %.4f ms'%d['e2e_pageable']['ms_per_step'])"
done
