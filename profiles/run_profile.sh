#!/bin/bash
# Profiling recipe behind profiles/*.md (run under gpurun, 1 GPU).  $1 = tag (e.g. r1)
#   1. bench line (never under a profiler)          -> gpurun_out/bench_$TAG.json
#   2. device timeline of one solve-loop iteration  -> gpurun_out/timeline_$TAG.txt
#   3. ncu launch list of bench.py (our kernels)    -> gpurun_out/launches_$TAG.csv
#   4. ncu --set full of the level-0 hot kernels    -> gpurun_out/prof_$TAG.ncu-rep
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
KREGEX='regex:sell_|dense_sym|reduce_partials|gather_system|scatter_'
python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 400 $OUT/bench_$TAG.err
python profiles/timeline.py > $OUT/timeline_$TAG.txt 2>&1
# launch list: skip the warm-up solve (first ~1500 of our launches), list 2 timed iterations
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 1500 -c 400 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launch_$TAG.log 2>&1
# full capture of the level-0 kernels (GS phases, residual, norm, restrict, prolong)
ncu --set full --clock-control none --import-source on -k 'regex:sell_gs_phase|sell_apply|sell_residual_norm' \
    -c 14 -o $OUT/prof_$TAG -f python profiles/kernel_probe.py --reps 1 > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT | tail -8
