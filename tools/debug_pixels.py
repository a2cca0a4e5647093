#!/usr/bin/env python3
"""Diagnoses pixels where the device render differs from the reference: renders a bench workload on both, lists the
differing pixels, replays them in the plain-C port (which must agree with the reference), and hands the port's path rays
to the device traversal to see whether a hit record or the shading differs.

  python tools/debug_pixels.py c2 4        # workload, spp
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "yocto-hair_b200")); sys.path.insert(0, str(ROOT / "oracle"))
import bench  # noqa: E402
import oracle  # noqa: E402
from yocto_hair_b200.pathtrace import DeviceScene, HIT_DTYPE, RAY_DTYPE  # noqa: E402
from yocto_hair_b200.scene import make_params  # noqa: E402

wl, spp = sys.argv[1], int(sys.argv[2])
scene, cam, res, _ = bench.build_workload(wl, True)
p = make_params(resolution=res)
ref, port = oracle.RefOracle(), oracle.PortOracle()
rs = ref.scene(scene)
sr = rs.state(cam, p)
ds = DeviceScene(scene).init_bvh().init_lights()
sd = ds.init_state(cam, p)
first = None
for k in range(1, spp + 1):
    sr.trace_samples(1); sd.trace_samples(1)
    a_r, a_d = sr.download()[1], sd.download()[1]
    diff = np.nonzero((a_r.view(np.uint32) != a_d.view(np.uint32)).any(axis=-1))
    print("after %d spp: %d differing pixels %s" % (k, diff[0].size, list(zip(diff[0][:6], diff[1][:6]))))
    if diff[0].size and first is None:
        first = (k, list(zip(diff[0], diff[1])))
if first is None:
    print("no difference"); sys.exit(0)
k, pixels = first
ps = port.scene(scene)
lib = port.lib
lib.yho_debug_pixel.restype = C.c_int
for (y, x) in pixels[:4]:
    sp = ps.state(cam, p)
    rays, hits = np.zeros(4096, RAY_DTYPE), np.zeros(4096, HIT_DTYPE)
    n = lib.yho_debug_pixel(C.c_void_p(sp.handle), int(x), int(y), k, rays.ctypes.data_as(C.c_void_p), hits.ctypes.data_as(C.c_void_p), 4096)
    a_p = sp.download()[1]
    print("pixel (y=%d, x=%d): port accum %s  ref %s  device %s; port == ref: %s" % (
        y, x, a_p[y, x], a_r[y, x], a_d[y, x], np.array_equal(a_p[y, x].view(np.uint32), a_r[y, x].view(np.uint32))))
    hd = ds.intersect(rays[:n])
    for q in range(n):
        same = (hits[q]["hit"] == hd[q]["hit"]) and (not hits[q]["hit"] or (hits[q]["object"] == hd[q]["object"] and hits[q]["element"] == hd[q]["element"]
                and hits[q]["distance"].view(np.uint32) == hd[q]["distance"].view(np.uint32) and np.array_equal(hits[q]["uv"].view(np.uint32), hd[q]["uv"].view(np.uint32))))
        print("  ray %2d o=%s d=%s  port hit=%d obj=%d el=%d t=%.9g uv=%s | dev hit=%d obj=%d el=%d t=%.9g uv=%s %s" % (
            q, rays[q]["o"], rays[q]["d"], hits[q]["hit"], hits[q]["object"], hits[q]["element"], hits[q]["distance"], hits[q]["uv"],
            hd[q]["hit"], hd[q]["object"], hd[q]["element"], hd[q]["distance"], hd[q]["uv"], "" if same else "  <-- DIFFERS"))
    sp.close()
