#!/bin/bash
# round-2 development check (run under gpurun): GPU tests, bench with / without the persistent
# smoother kernel, ogre timeline, MCF precompute stages, k = 3 timeline
OUT=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/r2_gpu_tests3.log 2>&1; tail -5 $OUT/r2_gpu_tests3.log
for gs in 1 0; do
  SMG_GS_STREAM=$gs timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/r2_bench3_gs$gs.json 2> $OUT/r2_bench3_gs$gs.err
  tail -c 300 $OUT/r2_bench3_gs$gs.err
  python - <<P
import json
d=json.load(open("$OUT/r2_bench3_gs$gs.json")); r=d["roofline"]
print("gs_stream=$gs value", d["value"], "e2e", d["e2e"]["value"], "cold frac", r["frac"], "warm", r["warm_pair"]["frac"], "iter", r["iteration"]["frac"], "insitu gs", r["in_situ"].get("l0_gs_frac"))
print({k: (round(v["ms"]*1e3,2), round(v["frac"],3)) for k,v in r["kernels"].items()})
P
done
timeout 200 python profiles/timeline.py --workload ogre > $OUT/r2_timeline_ogre.txt 2>&1; grep -E "levels|^total" $OUT/r2_timeline_ogre.txt; grep -A40 "exclusive time" $OUT/r2_timeline_ogre.txt | head -50
SMG_PRECOMPUTE_TIMING=1 timeout 300 python bench.py --workload mcf --flow-steps 2 --warmup 1 2>&1 | grep "numeric setup" | tail -14
timeout 200 python profiles/timeline.py --k 3 > $OUT/r2_timeline_k3.txt 2>&1; grep -A40 "exclusive time" $OUT/r2_timeline_k3.txt
