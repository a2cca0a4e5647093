#!/bin/bash
# tools/run_n.sh N : the bench on N GPUs of one box, as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err
echo "exit $?"; tail -2 gpurun_out/bench_c3_n$N.err; python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c3_n$N.json").read().strip().splitlines()[-1])
print("N=$N value %.1fM e2e %.1fM ms/step %.2f reduce_ms %s scaling %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d.get("reduce_ms"), d["scaling"]))
PY
