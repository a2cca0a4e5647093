#!/usr/bin/env python3
"""One traversal experiment on the bench workload: BVH knobs -> counters + times.

  python tools/exp_traverse.py c3 <max_split> <leaf_size> [spp]
Set YH_B200_LIB=.../libyocto_hair_b200_stats.so to get nodes / prims per ray.
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "yocto-hair_b200"))
import bench  # noqa: E402
from yocto_hair_b200.pathtrace import DeviceScene  # noqa: E402
from yocto_hair_b200.scene import make_params  # noqa: E402

wl, max_split, leaf = sys.argv[1], 0, 0
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
scene, cam, res, _ = bench.build_workload(wl, True)
t0 = time.time()
ds = DeviceScene(scene).init_bvh(max_split, leaf).init_lights()
build_s = time.time() - t0
bs = ds.bvh_stats()
st = ds.init_state(cam, make_params(resolution=res))
st.trace_samples(2)
s = st.trace_samples(spp)
print(json.dumps(dict(
    workload=wl, max_split=max_split, leaf=leaf, build_s=round(build_s, 3), refs=bs.num_references, nodes=bs.num_nodes,
    node_mb=round(bs.node_bytes / 1e6, 1), prim_mb=round(bs.prim_bytes / 1e6, 1),
    msamples_s=round(s.samples / s.device_ms / 1e3, 2), mrays_s=round(s.rays / s.device_ms / 1e3, 1),
    device_ms=round(s.device_ms, 2), traverse_ms=round(s.traverse_ms, 2), shade_ms=round(s.shade_ms, 2),
    nodes_per_ray=round(s.nodes_visited / max(1, s.rays), 2), prims_per_ray=round(s.prims_tested / max(1, s.rays), 2),
    iters=s.iterations)))
