#!/usr/bin/env python3
"""Times the set-up phases of the C ABI on the bench workload (GPU box)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "yocto-hair_b200"))
import bench
from yocto_hair_b200.pathtrace import DeviceScene
from yocto_hair_b200.scene import make_params
sc, cam, res, _ = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else "c3", True)
for rep in range(3):
    t = [time.perf_counter()]
    ds = DeviceScene(sc); t.append(time.perf_counter())
    ds.init_bvh(); t.append(time.perf_counter())
    ds.init_lights(); t.append(time.perf_counter())
    st = ds.init_state(cam, make_params(resolution=res)); t.append(time.perf_counter())
    st.trace_samples(1); t.append(time.perf_counter())
    st.close(); ds.close(); t.append(time.perf_counter())
    print("scene_create %.1f ms  init_bvh %.1f  init_lights %.1f  state_create %.1f  first spp %.1f  close %.1f" % tuple(1e3 * (b - a) for a, b in zip(t, t[1:])))
