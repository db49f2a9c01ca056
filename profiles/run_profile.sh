#!/bin/bash
# Profiling recipe used for profiles/ (run under gpurun, 1 GPU).  $1 = tag (e.g. r1a)
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
# 1. the bench line (never under a profiler)
python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 600 $OUT/bench_$TAG.err
# 2. launch list of the same command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launch_$TAG.log 2>&1
# 3. one full capture of the dominant kernels (fine-level GS phase, residual)
ncu --set full --clock-control none --import-source on -k regex:'sell_gs_phase|sell_apply|sell_residual_norm' \
    -s 40 -c 12 -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT
