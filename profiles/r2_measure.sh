#!/bin/bash
# Round-2 results table: every BASELINE config on one B200 with the
# reference's own sources timed beside it on the box's host CPU (1 thread: the reference path has no threading).
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 20 --warmup 3 > $OUT/final_sphere.json 2> $OUT/final_sphere.err
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/final_sphere_ref.json 2>> $OUT/final_sphere.err
for w in bunny ogre; do
  python bench.py --workload $w --steps 20 --warmup 3 --cpu-steps 20 > $OUT/final_$w.json 2> $OUT/final_$w.err
done
python bench.py --workload hilbert --steps 10 --warmup 3 --cpu-steps 3 > $OUT/final_hilbert.json 2> $OUT/final_hilbert.err
python bench.py --workload mcf --warmup 2 > $OUT/final_mcf.json 2> $OUT/final_mcf.err
python bench.py --workload mcf --impl reference --steps 2 > $OUT/final_mcf_ref.json 2>> $OUT/final_mcf.err
python bench.py --smoother wavefront --steps 10 --warmup 3 --no-cpu > $OUT/final_wavefront.json 2> $OUT/final_wavefront.err
for f in sphere sphere_ref bunny ogre hilbert mcf mcf_ref wavefront; do
python - <<P
import json
try:
    d=json.loads([l for l in open("$OUT/final_$f.json") if l.startswith("{")][-1])
    cb=d.get("cpu_baseline") or {}
    print("$f", "value %.1f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.1f" % d["e2e"]["value"], d["e2e"].get("step",""), "cpu", cb.get("value"), d.get("mcf",""))
except Exception as e:
    print("$f", "FAILED", e)
P
done
