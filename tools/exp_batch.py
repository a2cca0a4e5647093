#!/usr/bin/env python3
"""Steady-state traversal rate: one big batch of incoherent rays through yh_intersect_batch_dev."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "yocto-hair_b200"))
import bench
from yocto_hair_b200 import synth
from yocto_hair_b200.pathtrace import DeviceScene
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
scene, cam, res, _ = bench.build_workload("c3", True)
ds = DeviceScene(scene).init_bvh().init_lights()
rays = synth.random_rays(scene, n)
hits, ms = ds.intersect_timed(rays, repeats=3)
print("batch %d rays: %.3f ms = %.1f Mrays/s, hit fraction %.3f" % (n, ms, n / ms / 1e3, hits["hit"].mean()))
