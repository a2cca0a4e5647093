#!/usr/bin/env python3
"""Exploratory GPU-vs-oracle comparison (intersections + small renders)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "yocto-hair_b200"))
import numpy as np
import oracle
from yocto_hair_b200 import synth
from yocto_hair_b200.pathtrace import DeviceScene
from yocto_hair_b200.scene import make_params

o = oracle.best_oracle()
print("oracle:", o.kind)
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
res = int(sys.argv[2]) if len(sys.argv) > 2 else 128
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 4
for name, (sc, cam) in dict(block=synth.hair_block_scene(nseg), straight=synth.straight_hair_scene(nseg),
                            curly=synth.curly_hair_scene(2 * nseg)).items():
    rs = o.scene(sc)
    t = time.time(); ds = DeviceScene(sc).init_bvh().init_lights(); tb = time.time() - t
    bs = ds.bvh_stats()
    print(f"== {name}: build {tb*1e3:.1f} ms prims {bs.num_primitives} refs {bs.num_references} nodes {bs.num_nodes}")
    rays = np.concatenate([synth.random_rays(sc, 200000), synth.camera_rays(cam, res, res, 100000)])
    hr = rs.intersect(rays.view(np.float32).reshape(-1, 8), threads=8)
    hd, ms = ds.intersect_timed(rays)
    same_hit = hr["hit"] == hd["hit"]
    both = (hr["hit"] == 1) & (hd["hit"] == 1)
    ids = (hr["object"][both] == hd["object"][both]) & (hr["element"][both] == hd["element"][both])
    exact = ids & (hr["distance"][both].view(np.uint32) == hd["distance"][both].view(np.uint32)) & \
        (hr["uv"][both].view(np.uint32) == hd["uv"][both].view(np.uint32)).all(axis=1)
    print(f"  rays {len(rays)} hit frac {hr['hit'].mean():.3f}  hit-flag mismatches {np.count_nonzero(~same_hit)}  "
          f"id mismatches {np.count_nonzero(~ids)}  non-bit-exact uv/dist {np.count_nonzero(~exact)}  kernel {ms:.3f} ms "
          f"= {len(rays)/ms/1e3:.1f} Mrays/s")
    bad = np.nonzero(~same_hit)[0][:3]
    for b in bad: print("   mismatch", rays[b], hr[b], hd[b])
    bi = np.nonzero(both)[0][~ids]
    for b in bi[:16]:
        print("   id mismatch ray", b, "ref", hr[b]["object"], hr[b]["element"], repr(hr[b]["distance"]), hr[b]["uv"],
              "dev", hd[b]["object"], hd[b]["element"], repr(hd[b]["distance"]), hd[b]["uv"])
    p = make_params(resolution=res)
    st_r = rs.state(cam, p); tr = st_r.trace_samples(spp); r_r, a_r, s_r, rng_r = st_r.download()
    st_d = ds.init_state(cam, p); stats = st_d.trace_samples(spp); r_d, a_d, s_d = st_d.download(); rng_d = st_d.rng()
    eq = (a_r.view(np.uint32) == a_d.view(np.uint32)).all(axis=2)
    rng_eq = (rng_r == rng_d).all(axis=1)
    print(f"  render {st_d.width}x{st_d.height}x{spp}: identical pixels {eq.mean():.6f} rng-identical {rng_eq.mean():.6f} "
          f"max abs diff {np.abs(a_r-a_d).max():.3g}  mean ref {r_r[...,:3].mean():.5f} dev {r_d[...,:3].mean():.5f}")
    print(f"  cpu {tr:.3f}s ({st_d.width*st_d.height*spp/tr/1e6:.2f} Msamples/s, {o.hardware_concurrency()} threads)  gpu {stats.device_ms:.2f} ms "
          f"({stats.samples/stats.device_ms/1e3:.2f} Msamples/s) rays {stats.rays} light_rays {stats.light_rays} iters {stats.iterations}")
    if not eq.all():
        j, i = np.argwhere(~eq)[0]
        print("   first differing pixel", i, j, a_r[j, i], a_d[j, i], rng_r[j*st_d.width+i], rng_d[j*st_d.width+i])
