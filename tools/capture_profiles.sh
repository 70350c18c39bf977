#!/bin/bash
# ncu captures of the hot kernels at the benchmark sizes + the launch list of bench.py (run under gpurun).
set -x
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:ntt_pass_kernel -s 3 -c 3 -f -o gpurun_out/cap_lde python tools/profile_target.py lde 2 > gpurun_out/cap_lde.log 2>&1
timeout 300 $NCU -k regex:merkle_levels_kernel -s 4 -c 2  # last node kernel of build 1, then the leaf kernel of build 2 -f -o gpurun_out/cap_merkle python tools/profile_target.py merkle 2 > gpurun_out/cap_merkle.log 2>&1
timeout 300 $NCU -k regex:fri_fold_kernel -s 21 -c 1 -f -o gpurun_out/cap_fold python tools/profile_target.py fri 2 > gpurun_out/cap_fold.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/cap_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/cap_bench_under_ncu.log 2>&1
ls -la gpurun_out/
