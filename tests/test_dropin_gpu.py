"""The drop-in command line on a B200: host/_build/yscenetrace_b200 (reference CLI +
reference JSON/PLY/HDR loaders + the C++ host shim over the C ABI) renders a scene
with `--backend cuda` and with `--backend cpu` (the reference's own functions);
the saved .pfm images must be bit-identical."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from yocto_hair_b200 import scenefiles, synth

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
APP = ROOT / "yocto-hair_b200" / "host" / "_build" / "yscenetrace_b200"


def run_app(*args):
    run = subprocess.run([str(APP), *map(str, args)], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
    return run.stdout


@pytest.mark.parametrize("name,make,res", [
    ("curly_sky", lambda: synth.curly_hair_scene(30_000), 160),              # sky.hdr environment texture
    ("straight_lights", lambda: synth.straight_hair_scene(20_000), 128),     # sky + two area lights
    ("block", lambda: synth.hair_block_scene(20_000), 192),                  # diffuse sphere, 16:9
])
def test_cuda_backend_equals_cpu_backend(name, make, res, tmp_path):
    if not APP.exists():
        pytest.skip("host/_build/yscenetrace_b200 not built (needs /root/reference at build time)")
    scene, cam = make()
    js = scenefiles.write_scene(tmp_path, name, scene, cam)
    out = run_app("--backend", "cuda", "-s", 5, "-r", res, "--spp-batch", 3, "-o", tmp_path / "gpu.pfm", js)
    assert "backend cuda" in out
    run_app("--backend", "cpu", "-s", 5, "-r", res, "-o", tmp_path / "cpu.pfm", js)
    a, b = scenefiles.read_pfm(tmp_path / "gpu.pfm"), scenefiles.read_pfm(tmp_path / "cpu.pfm")
    assert a.shape == b.shape and a.shape[1] == res
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert a.mean() > 0.01


def test_save_batch_matches_one_sample_at_a_time(tmp_path):
    if not APP.exists():
        pytest.skip("host/_build/yscenetrace_b200 not built")
    scene, cam = synth.curly_hair_scene(10_000)
    js = scenefiles.write_scene(tmp_path, "batch", scene, cam)
    run_app("--backend", "cuda", "-s", 3, "-r", 96, "--save-batch", "-o", tmp_path / "g.pfm", js)
    run_app("--backend", "cpu", "-s", 3, "-r", 96, "--save-batch", "-o", tmp_path / "c.pfm", js)
    for k in range(3):  # <name>-s<k>.pfm after every pass, as the reference app names them
        a, b = scenefiles.read_pfm(tmp_path / f"g-s{k}.pfm"), scenefiles.read_pfm(tmp_path / f"c-s{k}.pfm")
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
