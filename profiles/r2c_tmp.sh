OUT=gpurun_out
timeout 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/bench_r2.json 2> $OUT/bench_r2.err; tail -c 200 $OUT/bench_r2.err
python -c "
import json
d=json.load(open('$OUT/bench_r2.json')); r=d['roofline']
print('value', d['value'], 'e2e', d['e2e']['value'], 'cold', r['frac'], 'warm', r['warm_pair']['frac'], 'insitu', r['in_situ'].get('l0_gs_frac'), 'iter', r['iteration']['frac'], 'cpu', d['cpu_baseline']['value'])
"
timeout 200 python profiles/timeline.py > $OUT/timeline_r2.txt 2>&1; grep -A30 "exclusive time" $OUT/timeline_r2.txt | grep -v "\.\(blob\|gather\|phases\|tail\)"
SMG_PROBE_DIST=0 SMG_NO_TMA=1 timeout 600 compute-sanitizer --tool synccheck --print-limit 20 --error-exitcode 9 python profiles/sanitize_probe.py > $OUT/sanitizer_synccheck_full.log 2>&1
rc=$?
(grep "kernels.cu:" $OUT/sanitizer_synccheck_full.log | sed 's/.*in kernels.cu/kernels.cu/' | sort | uniq -c; tail -3 $OUT/sanitizer_synccheck_full.log; echo "synccheck (SMG_NO_TMA=1) rc=$rc") > $OUT/sanitizer_synccheck_r2.log
rm -f $OUT/sanitizer_synccheck_full.log; cat $OUT/sanitizer_synccheck_r2.log
