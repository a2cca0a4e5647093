#!/bin/bash
for w in c3; do
  for g in 1 2 3; do
    YH_WAVE_GROUPS=$g python bench.py --workload $w --no-cpu-baseline --no-counter-baseline --steps 3 > gpurun_out/g_${g}_$w.json 2>/dev/null
    python - <<PY
import json
d=json.loads(open("gpurun_out/g_${g}_$w.json").read().strip().splitlines()[-1])
print("$w groups $g value %.1fM e2e %.1fM trav %.0f shade %.0f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["traverse_ms"], d["shade_ms"]))
PY
  done
done
for g in 1 2; do
  YH_WAVE_GROUPS=$g python bench.py --workload c2 --no-cpu-baseline --steps 3 > gpurun_out/gc_${g}_c2.json 2>/dev/null
  grep -o "n1_counter_mode_value[^,]*" gpurun_out/gc_${g}_c2.json
done
