#!/usr/bin/env python3
"""BVH builder comparison on a bench workload: quality 0 (device LBVH) vs 1 (host binned SAH).
  python tools/exp_quality.py c3 [spp]      (YH_B200_LIB=..._stats.so adds nodes / prims per ray)"""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "yocto-hair_b200"))
import bench
from yocto_hair_b200.pathtrace import DeviceScene
from yocto_hair_b200.scene import make_params
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
scene, cam, res, _ = bench.build_workload(wl, True)
for quality in (0, 1, 0, 1):
    ds = DeviceScene(scene)
    t0 = time.perf_counter(); ds.init_bvh(0, 0, quality); build_ms = 1e3 * (time.perf_counter() - t0)
    ds.init_lights()
    st = ds.init_state(cam, make_params(resolution=res))
    st.trace_samples(2)
    s = st.trace_samples(spp)
    print(json.dumps(dict(workload=wl, quality=quality, build_ms=round(build_ms, 1), msamples_s=round(s.samples / s.device_ms / 1e3, 2),
                          device_ms=round(s.device_ms, 2), nodes_per_ray=round(s.nodes_visited / max(1, s.rays), 2),
                          prims_per_ray=round(s.prims_tested / max(1, s.rays), 2))))
    st.close(); ds.close()
