#!/bin/bash
# ncu evidence of a round (one GPU): launch list of a short c3 bench run, --set full captures of the traversal kernel
# (c3: L2-resident tree; c5: the tree that leaves L2) and of the hair shading kernel.  Outputs -> gpurun_out/.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --spp-step 4 --no-cpu-baseline --no-counter-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 34 -c 1 -f -o gpurun_out/k_intersect $B > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_shade$' -s 61 -c 2 -f -o gpurun_out/k_shade $B > gpurun_out/ncu_full_shade.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 34 -c 1 -f -o gpurun_out/k_intersect_c5 $B --workload c5 > gpurun_out/ncu_full_c5.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_group_reduce -c 1 -f -o gpurun_out/k_group_reduce python - > gpurun_out/ncu_group.log 2>&1 <<'PY'
import sys
sys.path.insert(0, "yocto-hair_b200")
from yocto_hair_b200 import synth
from yocto_hair_b200.pathtrace import DeviceGroup
from yocto_hair_b200.scene import make_params
sc, cam = synth.curly_hair_scene(50_000)
g = DeviceGroup(sc, devices=[0, 0]).init_state(cam, make_params(resolution=1280), "tiles")
g.trace_samples(1); g.download()
PY
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_bench.log
