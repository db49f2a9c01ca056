#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/r2_gpu_tests4.log 2>&1; tail -4 $OUT/r2_gpu_tests4.log
timeout 200 python profiles/timeline.py --workload ogre > $OUT/r2_timeline_ogre2.txt 2>&1; grep -E "levels|^total" $OUT/r2_timeline_ogre2.txt; grep -A40 "exclusive time" $OUT/r2_timeline_ogre2.txt | head -30
SMG_PRECOMPUTE_TIMING=1 timeout 300 python bench.py --workload mcf --flow-steps 3 --warmup 1 2> $OUT/r2_mcf2.err | cut -c1-200; grep "numeric setup" $OUT/r2_mcf2.err | tail -7
for w in bunny ogre; do timeout 200 python bench.py --workload $w --steps 20 --warmup 3 > $OUT/r2b_$w.json 2> $OUT/r2b_$w.err; tail -c 300 $OUT/r2b_$w.err; python -c "
import json
d=json.load(open('$OUT/r2b_$w.json'))
print('$w', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_solve'], d['cpu_baseline']['value'] if d['cpu_baseline'] else None)
"; done
