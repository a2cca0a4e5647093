#!/bin/bash
# Quick GPU-box pass while iterating on a kernel: parity tests, then the headline bench without the CPU leg.
# usage: gpu_quick.sh [extra bench workloads...]   e.g. gpu_quick.sh c2 c5
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for w in c3 "$@"; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/quick_$w.json 2> gpurun_out/quick_$w.err; echo "bench $w exit $?"
  tail -3 gpurun_out/quick_$w.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/quick_$w.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$w", "samples/s %.1fM e2e %.1fM mrays %.0f trav_ms %.1f shade_ms %.1f nodes/ray %.1f prims/ray %.2f frac %.3f" % (
        d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["mrays_per_sec"], d["traverse_ms"], d["shade_ms"],
        r.get("nodes_per_ray", 0), r.get("prims_per_ray", 0), r.get("frac") or 0))
except Exception as e:
    print("no bench line:", e)
PY
done
