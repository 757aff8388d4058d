#!/bin/bash
OUT=gpurun_out/${1:-tl}; mkdir -p $OUT
timeout 200 python scripts/lk_timeline.py 2>&1 | tee $OUT/timeline_win21.log
timeout 200 python scripts/lk_timeline.py win31 2>&1 | tee $OUT/timeline_win31.log
KLT_LK_WPP=2 timeout 200 python scripts/lk_timeline.py 2>&1 | tee $OUT/timeline_win21_wpp2.log
timeout 200 python scripts/lk_cycles.py 2>&1 | tee $OUT/cycles_win21.log
