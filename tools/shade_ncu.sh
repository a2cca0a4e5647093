mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_shade$' -s 61 -c 2 -f -o gpurun_out/k_shade python bench.py --steps 1 --warmup 1 --spp-step 4 --no-cpu-baseline > gpurun_out/ncu_full_shade.log 2>&1
tail -3 gpurun_out/ncu_full_shade.log
