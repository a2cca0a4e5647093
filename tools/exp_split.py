#!/usr/bin/env python3
"""Reference-splitting sweep on a bench workload: max_split -> refs, samples/s (and, with the stats library, visits per ray).
  python tools/exp_split.py c3 "1 2 4 6 8" [spp]"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "yocto-hair_b200"))
import bench
from yocto_hair_b200.pathtrace import DeviceScene
from yocto_hair_b200.scene import make_params
wl = sys.argv[1]
splits = [int(x) for x in sys.argv[2].split()]
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 16
scene, cam, res, _ = bench.build_workload(wl, True)
for ms in splits:
    ds = DeviceScene(scene).init_bvh(ms, 0).init_lights()
    bs = ds.bvh_stats()
    st = ds.init_state(cam, make_params(resolution=res))
    st.trace_samples(2)
    s = st.trace_samples(spp)
    print(json.dumps(dict(workload=wl, max_split=ms, refs=bs.num_references, msamples_s=round(s.samples / s.device_ms / 1e3, 2),
                          nodes_per_ray=round(s.nodes_visited / max(1, s.rays), 2), prims_per_ray=round(s.prims_tested / max(1, s.rays), 2))))
    st.close(); ds.close()
