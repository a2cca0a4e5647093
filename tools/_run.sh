#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
bash tools/gpu_bench_all.sh
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err
B="python bench.py --steps 1 --warmup 1 --spp-step 4 --no-cpu-baseline --no-counter-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 34 -c 1 -f -o gpurun_out/k_intersect $B > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_shade$' -s 61 -c 2 -f -o gpurun_out/k_shade $B > gpurun_out/ncu_full_shade.log 2>&1
ls -la gpurun_out/*.ncu-rep
