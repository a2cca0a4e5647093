"""Drop-in plumbing on CPU: the harness writes a synthetic hair scene in the
reference's own on-disk formats (scene JSON + binary PLY polylines), the drop-in
command line host/_build/yscenetrace_b200 loads it through the reference's
unmodified yocto_sceneio and renders it with `--backend cpu` (the reference's
own functions).  The image must equal, bit for bit, the render of the same
scene handed to the reference IN MEMORY — which pins the file writers, the
io -> ptr scene conversion of the app and the flattened C-ABI description to
each other.  (`--backend cuda` is compared with `--backend cpu` in
tests/test_dropin_gpu.py.)"""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle as orc
from yocto_hair_b200 import scenefiles, synth
from yocto_hair_b200.scene import make_params

ROOT = Path(__file__).resolve().parent.parent
APP = ROOT / "yocto-hair_b200" / "host" / "_build" / "yscenetrace_b200"


def need_app():
    if not APP.exists():
        pytest.skip("host/_build/yscenetrace_b200 not built (needs /root/reference at build time)")
    if not orc.REF_LIB.exists():
        pytest.skip("oracle/_ref/libyh_ref.so not built")


@pytest.mark.parametrize("name,make", [
    ("curly", lambda: synth.curly_hair_scene(3_000, sky=False)),
    ("straight", lambda: synth.straight_hair_scene(2_000, sky=False)),   # + two area lights (triangle quads)
    ("block", lambda: synth.hair_block_scene(2_000)),                    # + diffuse sphere, 16:9 camera
])
def test_files_roundtrip_through_the_reference_loader(name, make, tmp_path):
    need_app()
    scene, cam = make()
    js = scenefiles.write_scene(tmp_path, name, scene, cam)
    out = tmp_path / "cpu.pfm"
    run = subprocess.run([str(APP), "--backend", "cpu", "-s", "2", "-r", "72", "-o", str(out), str(js)],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
    img = scenefiles.read_pfm(out)
    ref = orc.RefOracle()
    rs = ref.scene(scene)
    st = rs.state(cam, make_params(resolution=72))
    st.trace_samples(2)
    render = st.download()[0]
    assert img.shape[:2] == render.shape[:2]
    assert np.array_equal(img.view(np.uint32), np.ascontiguousarray(render[..., :3]).view(np.uint32))
    st.close(); rs.close()


def test_app_rejects_out_of_scope_scenes_loudly(tmp_path):
    need_app()
    scene, cam = synth.curly_hair_scene(500, sky=False)
    scene.materials[0].specular = 0.5
    js = scenefiles.write_scene(tmp_path, "spec", scene, cam)
    run = subprocess.run([str(APP), "--backend", "cuda", "-s", "1", "-r", "32", "-o", str(tmp_path / "x.pfm"), str(js)],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode != 0  # no device here -> YH_ERR_NO_DEVICE; on a GPU box -> YH_ERR_UNSUPPORTED; never a silent CPU render
    assert not (tmp_path / "x.pfm").exists()
