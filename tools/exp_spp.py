#!/usr/bin/env python3
"""samples/s as a function of the samples per pixel handed to one yh_trace_samples call (tail of the wavefront).
  python tools/exp_spp.py c3 "4 8 16 32 64 128" """
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "yocto-hair_b200"))
import bench
from yocto_hair_b200.pathtrace import DeviceScene
from yocto_hair_b200.scene import make_params
wl = sys.argv[1]
scene, cam, res, _ = bench.build_workload(wl, True)
ds = DeviceScene(scene).init_bvh().init_lights()
st = ds.init_state(cam, make_params(resolution=res))
st.trace_samples(4)
for spp in [int(x) for x in sys.argv[2].split()]:
    s = st.trace_samples(spp)
    print(json.dumps(dict(workload=wl, spp=spp, msamples_s=round(s.samples / s.device_ms / 1e3, 2), device_ms=round(s.device_ms, 2), iters=s.iterations)))
