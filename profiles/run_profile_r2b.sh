#!/bin/bash
# second half of profiles/run_profile_r2.sh: the ncu passes need SMG_HOST_LOOP=1 (ncu cannot profile
# kernel nodes of a graph that contains conditional nodes, i.e. the device-side solve loop)
set -u
TAG=r2
OUT=gpurun_out
mkdir -p $OUT
export SMG_HOST_LOOP=1
KREGEX='regex:sell_|patch_kernel|dense_sym|gather_system|scatter_|solve_decide|halo_exchange'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 600 -c 300 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none -k 'regex:sell_gs_phase_multi|sell_apply_short_kernel' -s 60 -c 26 -o $OUT/prof_l0_$TAG -f \
    python profiles/kernel_probe.py --reps 1 > $OUT/ncu_full_l0_$TAG.log 2>&1
ncu -i $OUT/prof_l0_$TAG.ncu-rep --page raw --csv > $OUT/prof_l0_$TAG.csv 2>/dev/null; rm -f $OUT/prof_l0_$TAG.ncu-rep
ncu --set full --clock-control none -k 'regex:patch_kernel|dense_sym|sell_gs_phase_kernel' -c 24 -o $OUT/prof_small_$TAG -f \
    python profiles/kernel_probe.py --reps 1 --kernels residual > $OUT/ncu_full_small_$TAG.log 2>&1
ncu -i $OUT/prof_small_$TAG.ncu-rep --page raw --csv > $OUT/prof_small_$TAG.csv 2>/dev/null; rm -f $OUT/prof_small_$TAG.ncu-rep
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    -k 'regex:sell_gs_phase' -s 64 -c 24 --csv --log-file $OUT/traffic_nocc_$TAG.csv \
    python profiles/kernel_probe.py --reps 1 --kernels relax_sweep,relax_pre > $OUT/ncu_traffic_$TAG.log 2>&1
unset SMG_HOST_LOOP
SMG_NO_TMA=1 timeout 600 compute-sanitizer --tool synccheck --print-limit 2000 --error-exitcode 9 python profiles/sanitize_probe.py > $OUT/sanitizer_synccheck_full.log 2>&1
rc=$?
grep "kernels.cu:" $OUT/sanitizer_synccheck_full.log | sed 's/.*in kernels.cu/kernels.cu/' | sort | uniq -c > $OUT/sanitizer_synccheck_$TAG.log
tail -3 $OUT/sanitizer_synccheck_full.log >> $OUT/sanitizer_synccheck_$TAG.log
echo "synccheck (SMG_NO_TMA=1) rc=$rc" >> $OUT/sanitizer_synccheck_$TAG.log
rm -f $OUT/sanitizer_synccheck_full.log
cat $OUT/sanitizer_synccheck_$TAG.log
wc -c $OUT/prof_l0_$TAG.csv $OUT/prof_small_$TAG.csv $OUT/traffic_nocc_$TAG.csv $OUT/launches_$TAG.csv
du -sh $OUT
