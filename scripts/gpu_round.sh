#!/bin/bash
# One GPU call: parity tests, bench (both arms), ncu launch list of the bench command, full ncu captures of the kernels.
# Usage (under gpurun): bash scripts/gpu_round.sh <tag>     then here: python scripts/summarize_profiles.py gpurun_out/<tag> profiles/r02
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
python bench.py --impl reference --steps 200 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
tail -c 600 $OUT/bench.err
python bench.py --steps 20 --warmup 5 > $OUT/bench_steps20.json 2>> $OUT/bench.err; echo "bench --steps 20 exit $?"
python bench.py --workload kitti_ref_params --steps 200 --no-detection > $OUT/bench_win31.json 2>> $OUT/bench.err
python bench.py --workload malaga_seq > $OUT/bench_malaga_seq.json 2>> $OUT/bench.err; echo "malaga_seq exit $?"
# launch list of the same bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 466 -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-detection --no-sharded-batch > $OUT/bench_under_ncu.log 2>&1
# full captures: batched pyramid (level 0->1 of 310 KITTI images and the one-launch build), single-pair LK, batched LK
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pyr_ -c 4 -o $OUT/prof_pyr -f \
    python scripts/prof_target.py pyr > $OUT/prof_pyr.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lk_fast -s 1 -c 2 -o $OUT/prof_lk -f \
    python scripts/prof_target.py lk > $OUT/prof_lk.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lk_fast -c 1 -o $OUT/prof_lk_batch -f \
    python scripts/prof_target.py batch > $OUT/prof_lk_batch.log 2>&1
# detection step: launch list of goodFeaturesToTrack calls, full capture of the two running-sum kernels
timeout 300 python scripts/corners_time.py > $OUT/corners_time.log 2>&1
REPS=5 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/corners_launches.csv \
    python scripts/corners_time.py > /dev/null 2>&1
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 2 -o $OUT/prof_corners -f \
    python scripts/corners_time.py > $OUT/prof_corners.log 2>&1
ls -la $OUT
