#!/bin/bash
# Every BASELINE workload through bench.py on one GPU (N = 1): gpurun_out/bench_<workload>.json.  The default
# workload (c3) carries the CPU baseline; the others skip it unless CPU_ALL=1.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for w in ${WLS:-c3 c2 c1 c1sq c4 c5}; do
  extra="--no-cpu-baseline"
  if [ "$w" = c3 ] || [ -n "$CPU_ALL" ]; then extra=""; fi
  timeout 900 python bench.py --workload $w $extra ${BENCH_ARGS:-} > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"
  tail -2 gpurun_out/bench_$w.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$w", "samples/s %.1fM e2e %.1fM (1spp/call %.1fM) mrays %.0f trav_ms %.1f shade_ms %.1f nodes/ray %.1f prims/ray %.2f frac %.3f cpu %s" % (
        d["value"] / 1e6, d["e2e"]["value"] / 1e6, ((d["e2e"].get("per_sample_call") or {}).get("value") or 0) / 1e6, d["mrays_per_sec"],
        d["traverse_ms"], d["shade_ms"], r.get("nodes_per_ray", 0), r.get("prims_per_ray", 0), r.get("frac") or 0,
        (d.get("cpu_baseline") or {}).get("value")))
except Exception as e:
    print("no bench line:", e)
PY
done
