#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --spp-step 4 --no-cpu-baseline --no-counter-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 34 -c 1 -f -o gpurun_out/k_intersect $B > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_shade$' -s 61 -c 2 -f -o gpurun_out/k_shade $B > gpurun_out/ncu_full_shade.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_bench.log
