#!/bin/bash
# Multi-GPU pass on an N-GPU box: bench.py through torch.distributed.run at 1 / 2 / 4 / N ranks (as the driver launches
# it), the one-process yh_group path (--group), C5 at N ranks, and the group tests on real peers.
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | head -8 > gpurun_out/scale_gpus.txt
timeout 600 python -m pytest tests/test_group_gpu.py tests/test_dropin_gpu.py -x -q -k "group or two_gpus or nccl" 2>&1 | tail -4
run() { # n workload extra...
  n=$1; w=$2; shift 2
  if [ $n = 1 ]; then timeout 900 python bench.py --gpus 1 --workload $w --no-cpu-baseline "$@"
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --workload $w "$@"; fi
}
for n in 1 2 4 $N; do
  [ $n -gt $N ] && continue
  run $n c3 > gpurun_out/scale_c3_n$n.json 2> gpurun_out/scale_c3_n$n.err; echo "c3 n=$n exit $?"
done
run $N c3 --shard tiles > gpurun_out/scale_c3_tiles_n$N.json 2> gpurun_out/scale_c3_tiles_n$N.err; echo "c3 tiles n=$N exit $?"
run $N c5 > gpurun_out/scale_c5_n$N.json 2> gpurun_out/scale_c5_n$N.err; echo "c5 n=$N exit $?"
timeout 900 python bench.py --gpus $N --group > gpurun_out/scale_c3_group_n$N.json 2> gpurun_out/scale_c3_group_n$N.err; echo "c3 group n=$N exit $?"
timeout 900 python bench.py --gpus $N --group --shard tiles > gpurun_out/scale_c3_group_tiles_n$N.json 2> gpurun_out/scale_c3_group_tiles_n$N.err; echo "c3 group tiles n=$N exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "n=%d value %.1fM e2e %.1fM reduce_ms %.2f setup_ms %.1f" % (d["n_gpus"], d["value"] / 1e6, d["e2e"]["value"] / 1e6, d.get("reduce_ms") or 0, d["e2e"].get("setup_ms") or 0), d["config"].get("parallelism"))
    except Exception as e:
        print(f, "no line:", e)
PY
