#!/bin/bash
# Round-2 profiler captures of the headline search step (run on the B200 box, one GPU):
#   1. the command exits 0 without ncu first;
#   2. launch list (gpu__time_duration.sum per launch) of the same command -> gpurun_out/r2_launches_bench_10m.csv;
#   3. one `--set full` capture of the scan kernel -> gpurun_out/prof_r2_scan_10m.ncu-rep (summarised into profiles/ by
#      scripts/summarise_ncu.sh on the build machine).
# Under ncu kernels are serialised and cold-cache: compare SHARES with the bench's own event times, not absolutes.
set -x
export FRS_BENCH_NO_CPU=1
CMD="python bench.py --steps 5 --warmup 3 --no-secondary"
$CMD > gpurun_out/r2_profile_plain.log 2>&1 || { tail -20 gpurun_out/r2_profile_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_kernel|merge_kernel|prep_queries|merge_shards" -c 400 --csv --log-file gpurun_out/r2_launches_bench_10m.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 6 -c 1 -f -o gpurun_out/prof_r2_scan_10m $CMD > gpurun_out/r2_ncu_full.log 2>&1
tail -3 gpurun_out/r2_ncu_full.log
ls -la gpurun_out/prof_r2_scan_10m.ncu-rep
