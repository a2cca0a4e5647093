import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "yocto-hair_b200"))
from yocto_hair_b200 import synth
from yocto_hair_b200.pathtrace import DeviceScene
from yocto_hair_b200.scene import make_params
def check(tag, scene, cam, res, spp, one_by_one=False):
    ds = DeviceScene(scene).init_bvh().init_lights()
    out = []
    for rep in range(3):
        st = ds.init_state(cam, make_params(resolution=res))
        if one_by_one and rep == 0:
            rays = sum(st.trace_samples(1).rays for _ in range(spp))
        else:
            rays = st.trace_samples(spp).rays
        a = st.download()[1]
        out.append((rays, a.view(np.uint32).copy()))
        st.close()
    same = [np.array_equal(out[0][1], o[1]) for o in out[1:]]
    nd = [int((out[0][1] != o[1]).any(axis=-1).sum()) for o in out[1:]]
    print(tag, "%dx%d" % (a.shape[1], a.shape[0]), "rays", [o[0] for o in out], "identical", same, "diff px", nd, flush=True)
    ds.close()

sc, cam = synth.hair_block_scene(200_000)
check("c1 4spp (rep0 = 4 x 1spp)", sc, cam, 720, 4, True)
check("c1 4spp", sc, cam, 720, 4)
check("c1 res512", sc, cam, 512, 4)
check("c1 res1024", sc, cam, 1024, 4)
sc2, cam2 = synth.hair_block_scene(200_000, aspect=1.0)
check("c1sq", sc2, cam2, 720, 4)
sc3, cam3 = synth.hair_block_scene(24_000)
check("block24k", sc3, cam3, 720, 4)
sc4, cam4 = synth.curly_hair_scene(200_000)
check("curly200k", sc4, cam4, 720, 4)
