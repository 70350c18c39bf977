#!/bin/bash
# ncu captures of the hot kernels at the benchmark sizes + the launch list of bench.py (run under gpurun).
#   bash tools/capture_profiles.sh [tag]      -> gpurun_out/<tag>_*.{ncu-rep,csv,log}
TAG=${1:-cap}
set -x
NCU="ncu --set full --clock-control none --import-source on"
# the three pass kernels of one coset LDE 2^24 x 8 (second LDE of the process: tables are built by the first)
timeout 300 $NCU -k regex:ntt_pass_kernel -s 3 -c 3 -f -o gpurun_out/${TAG}_lde python tools/profile_target.py lde 2 > gpurun_out/${TAG}_lde.log 2>&1
# Merkle: last node kernel of build 1, then the leaf kernel of build 2
timeout 300 $NCU -k regex:merkle_levels_kernel -s 4 -c 2 -f -o gpurun_out/${TAG}_merkle python tools/profile_target.py merkle 2 > gpurun_out/${TAG}_merkle.log 2>&1
# first FRI fold (2^24 -> 2^23) of the second chain
timeout 300 $NCU -k regex:fri_fold_kernel -s 21 -c 1 -f -o gpurun_out/${TAG}_fold python tools/profile_target.py fri 2 > gpurun_out/${TAG}_fold.log 2>&1
for k in lde merkle fold; do
  ncu -i gpurun_out/${TAG}_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_$k.raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$k.ncu-rep --page details > gpurun_out/${TAG}_ncu_$k.details.txt 2>/dev/null
done
# launch list of a whole bench run (device times are cold-cache and serialised: compare shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep --no-fib --no-fields > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
