#!/bin/bash
# One GPU-box pass: parity tests, headline bench, ncu launch list, one full capture of the traversal kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench exit $?" >> gpurun_out/bench_c3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --spp-step 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 34 -c 1 -f -o gpurun_out/k_intersect \
   python bench.py --steps 1 --warmup 1 --spp-step 4 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_shade$' -s 61 -c 2 -f -o gpurun_out/k_shade \
   python bench.py --steps 1 --warmup 1 --spp-step 4 --no-cpu-baseline > gpurun_out/ncu_full_shade.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_c3.json
