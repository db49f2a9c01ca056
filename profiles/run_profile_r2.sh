#!/bin/bash
# Profiling recipe behind profiles/r2_*.md (run under gpurun, 1 GPU).  Keeps gpurun_out small:
# the .ncu-rep files are converted to CSV on the box and deleted.
#   1. bench line (never under a profiler)            -> gpurun_out/bench_r2.json
#   2. device timeline of one solve-loop iteration    -> gpurun_out/timeline_r2.txt
#   3. ncu launch list of bench.py (our kernels)      -> gpurun_out/launches_r2.csv
#   4. ncu --set full: level-0 kernels, then the patch / coarse kernels of one V-cycle
#   5. DRAM traffic of the smoother without ncu's cache flush (sweeps share L2 as in production)
#   6. compute-sanitizer: memcheck (everything on), racecheck / synccheck with SMG_NO_TMA=1 (the
#      tools crash on cp.async.bulk + mbarrier kernels), racecheck also without PDL
set -u
TAG=r2
OUT=gpurun_out
mkdir -p $OUT
KREGEX='regex:sell_|patch_kernel|dense_sym|gather_system|scatter_|solve_decide|halo_exchange'
python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 300 $OUT/bench_$TAG.err
python profiles/timeline.py > $OUT/timeline_$TAG.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 600 -c 300 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launch_$TAG.log 2>&1
# level-0 kernels (2 rows per thread variants) as kernel_probe launches them after its 2-iteration solve
ncu --set full --clock-control none -k 'regex:sell_gs_phase_multi|sell_apply_short_kernel' -s 60 -c 26 -o $OUT/prof_l0_$TAG -f \
    python profiles/kernel_probe.py --reps 1 > $OUT/ncu_full_l0_$TAG.log 2>&1
ncu -i $OUT/prof_l0_$TAG.ncu-rep --page raw --csv > $OUT/prof_l0_$TAG.csv 2>/dev/null; rm -f $OUT/prof_l0_$TAG.ncu-rep
# patch kernels + coarse solve + level-1 phase kernels of the first V-cycle
ncu --set full --clock-control none -k 'regex:patch_kernel|dense_sym|sell_gs_phase_kernel' -c 24 -o $OUT/prof_small_$TAG -f \
    python profiles/kernel_probe.py --reps 1 --kernels residual > $OUT/ncu_full_small_$TAG.log 2>&1
ncu -i $OUT/prof_small_$TAG.ncu-rep --page raw --csv > $OUT/prof_small_$TAG.csv 2>/dev/null; rm -f $OUT/prof_small_$TAG.ncu-rep
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    -k 'regex:sell_gs_phase' -s 64 -c 24 --csv --log-file $OUT/traffic_nocc_$TAG.csv \
    python profiles/kernel_probe.py --reps 1 --kernels relax_sweep,relax_pre > $OUT/ncu_traffic_$TAG.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/sanitize_probe.py > $OUT/sanitizer_memcheck_$TAG.log 2>&1
echo "memcheck rc=$?" >> $OUT/sanitizer_memcheck_$TAG.log
SMG_NO_TMA=1 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/sanitize_probe.py > $OUT/sanitizer_racecheck_$TAG.log 2>&1
echo "racecheck (SMG_NO_TMA=1, PDL on) rc=$?" >> $OUT/sanitizer_racecheck_$TAG.log
SMG_NO_TMA=1 SMG_NO_PDL=1 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/sanitize_probe.py > $OUT/sanitizer_racecheck_nopdl_$TAG.log 2>&1
echo "racecheck (SMG_NO_TMA=1, SMG_NO_PDL=1) rc=$?" >> $OUT/sanitizer_racecheck_nopdl_$TAG.log
SMG_NO_TMA=1 timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python profiles/sanitize_probe.py > $OUT/sanitizer_synccheck_$TAG.log 2>&1
echo "synccheck (SMG_NO_TMA=1) rc=$?" >> $OUT/sanitizer_synccheck_$TAG.log
for f in $OUT/sanitizer_*_$TAG.log; do echo "== $f"; tail -3 $f; done
du -sh $OUT; ls -la $OUT | tail -16
