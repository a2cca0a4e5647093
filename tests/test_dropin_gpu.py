"""The drop-in command line on a B200: host/_build/yscenetrace_b200 (reference CLI +
reference JSON/PLY/HDR loaders + the C++ host shim over the C ABI; CUDA only, no CPU
path) renders a scene, and the TEST tool host/_build/yscenetrace_ab (same source with
-DYH_AB_TOOL) renders it with `--backend cpu` (the reference's own functions); the
saved .pfm images must be bit-identical."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from yocto_hair_b200 import scenefiles, synth

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
APP = ROOT / "yocto-hair_b200" / "host" / "_build" / "yscenetrace_b200"
AB = ROOT / "yocto-hair_b200" / "host" / "_build" / "yscenetrace_ab"


def run_app(*args):
    """`--backend cuda ...` runs the product binary (which has no such flag: it is CUDA only),
    `--backend cpu ...` the A/B test tool."""
    args = [str(a) for a in args]
    exe = APP
    if args[:1] == ["--backend"]:
        exe = AB if args[1] == "cpu" else APP
        args = args[2:] if args[1] == "cuda" else args
    run = subprocess.run([str(exe), *args], capture_output=True, text=True, timeout=900)
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


def test_materials_and_byte_textures_through_the_app(tmp_path):
    """The whole material model through the files and the shim: surface lobes, a refractive scattering sphere, and
    BYTE textures (.png: sRGB-decoded in the colour slot, linear in the roughness / opacity slots - the shim uploads
    the two expansions, yocto_pathtrace_b200.cpp) with the reference's own loaders on both sides."""
    if not APP.exists():
        pytest.skip("host/_build/yscenetrace_b200 not built")
    import json
    from PIL import Image
    scene, cam = synth.surface_lobes_scene(3_000)
    vol, _ = synth.surface_lobes_scene(500, volumes=True)
    wax = vol.materials[2 + list(synth.VOLUME_MATERIALS).index("wax")]
    scene.materials.append(wax)
    from yocto_hair_b200.scene import Object
    scene.objects.append(Object(shape=1, material=len(scene.materials) - 1, frame=(1, 0, 0, 0, 1, 0, 0, 0, 1, 0.4, 0.38, 1.6)))
    js = scenefiles.write_scene(tmp_path, "materials", scene, cam)
    rng = np.random.default_rng(9)
    Image.fromarray(rng.integers(30, 255, (16, 24, 3), dtype=np.uint8)).save(tmp_path / "textures" / "checker.png")
    Image.fromarray(rng.integers(0, 255, (8, 8), dtype=np.uint8)).save(tmp_path / "textures" / "gray.png")   # scalar (1-channel) file
    d = json.loads(Path(js).read_text())
    d["materials"]["mat001"]["color_tex"] = "checker"        # floor: sRGB reading
    d["materials"]["mat002"]["roughness_tex"] = "checker"    # the SAME file read as linear
    d["materials"]["mat002"]["color_tex"] = "checker"
    d["materials"]["mat004"]["roughness_tex"] = "gray"
    d["materials"]["mat010"]["opacity_tex"] = "gray"
    d["materials"]["mat005"]["normal_tex"] = "checker"       # byte normal map on the mirror sphere (linear reading)
    Path(js).write_text(json.dumps(d))
    out = run_app("--backend", "cuda", "-s", 4, "-r", 160, "--spp-batch", 2, "-o", tmp_path / "gpu.pfm", js)
    assert "backend cuda" in out
    run_app("--backend", "cpu", "-s", 4, "-r", 160, "-o", tmp_path / "cpu.pfm", js)
    a, b = scenefiles.read_pfm(tmp_path / "gpu.pfm"), scenefiles.read_pfm(tmp_path / "cpu.pfm")
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert a.mean() > 0.01


CUBE_OBJ = """v -1 -1 -1\nv 1 -1 -1\nv 1 1 -1\nv -1 1 -1\nv -1 -1 1\nv 1 -1 1\nv 1 1 1\nv -1 1 1
vt 0 0\nvt 1 0\nvt 1 1\nvt 0 1
f 1/1 4/2 3/3 2/4\nf 5/1 6/2 7/3 8/4\nf 1/1 2/2 6/3 5/4\nf 2/1 3/2 7/3 6/4\nf 3/1 4/2 8/3 7/4\nf 4/1 1/2 5/3 8/4
"""


def write_subdiv_scene(tmp_path):
    """A hair block next to a Catmull-Clark cube (subdivs/<name>.obj, material.subdivisions) with a displacement map."""
    import json
    from PIL import Image
    scene, cam = synth.hair_block_scene(3_000)
    js = scenefiles.write_scene(tmp_path, "subdiv", scene, cam)
    (tmp_path / "subdivs").mkdir()
    (tmp_path / "subdivs" / "cube.obj").write_text(CUBE_OBJ)
    rng = np.random.default_rng(4)
    Image.fromarray(rng.integers(0, 255, (8, 8), dtype=np.uint8)).save(tmp_path / "textures" / "bumps.png")
    d = json.loads(Path(js).read_text())
    d["materials"]["clay"] = {"color": [0.7, 0.5, 0.4], "specular": 0.5, "roughness": 0.4, "subdivisions": 2,
                              "displacement": 0.08, "displacement_tex": "bumps"}
    d["objects"]["zcube"] = {"subdiv": "cube", "material": "clay", "frame": [0.3, 0, 0, 0, 0.3, 0, 0, 0, 0.3, 0.2, 0.35, 1.8]}
    Path(js).write_text(json.dumps(d))
    return js


def test_subdivision_surfaces_through_the_app(tmp_path):
    """Subdivision + displacement stay the reference's own host code (init_subdivs runs before either backend builds
    its BVH, as in apps/yscenetrace/yscenetrace.cpp:243); the tesselated shape then renders bit-identically."""
    if not APP.exists():
        pytest.skip("host/_build/yscenetrace_b200 not built")
    js = write_subdiv_scene(tmp_path)
    run_app("--backend", "cuda", "-s", 4, "-r", 160, "-o", tmp_path / "gpu.pfm", js)
    run_app("--backend", "cpu", "-s", 4, "-r", 160, "-o", tmp_path / "cpu.pfm", js)
    a, b = scenefiles.read_pfm(tmp_path / "gpu.pfm"), scenefiles.read_pfm(tmp_path / "cpu.pfm")
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_interactive_session_headless(tmp_path):
    """the interactive consumer's use of the boundary (apps/ysceneitraces/ysceneitraces.cpp:255-302)
    without the window - preview at resolution / pratio, a worker thread adding one sample per pass under a stop
    flag, a camera edit that resets the state, a render to completion.  The preview and the post-edit image are
    bit-identical between the backends; on the device the interrupted state equals a fresh render of the passes
    that completed (the app checks it and fails otherwise)."""
    if not APP.exists():
        pytest.skip("host/_build/yscenetrace_b200 not built")
    scene, cam = synth.straight_hair_scene(12_000)
    js = scenefiles.write_scene(tmp_path, "session", scene, cam)
    outs = {}
    for backend in ("cuda", "cpu"):
        d = tmp_path / backend
        d.mkdir()
        outs[backend] = run_app("--backend", backend, "-s", 5, "-r", 128, "--interactive-check", d, "--stop-after", 2, js)
    assert "interrupted state equals a fresh" in outs["cuda"]
    for name in ("preview.pfm", "edited.pfm"):
        a, b = scenefiles.read_pfm(tmp_path / "cuda" / name), scenefiles.read_pfm(tmp_path / "cpu" / name)
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), name
    assert scenefiles.read_pfm(tmp_path / "cuda" / "preview.pfm").shape[1] == 128 // 8  # params.pratio = 8


def test_app_on_two_gpus_is_bit_identical(tmp_path):
    """`--gpus 2`: one process, two replicas, tile shards summed on the first GPU over NVLink (yh_group_*) - the image
    equals the one-GPU image bit for bit; `--shard samples` renders the same scene in counter mode."""
    import ctypes as C
    from yocto_hair_b200 import _capi as capi
    if not APP.exists():
        pytest.skip("host/_build/yscenetrace_b200 not built")
    have = C.c_int()
    capi.check(capi.load().yh_device_count(C.byref(have)))
    if have.value < 2:
        pytest.skip("needs two GPUs")
    scene, cam = synth.curly_hair_scene(30_000)
    js = scenefiles.write_scene(tmp_path, "two", scene, cam)
    run_app("-s", 4, "-r", 160, "-o", tmp_path / "one.pfm", js)
    run_app("--gpus", 2, "-s", 4, "-r", 160, "-o", tmp_path / "two.pfm", js)
    a, b = scenefiles.read_pfm(tmp_path / "one.pfm"), scenefiles.read_pfm(tmp_path / "two.pfm")
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    run_app("--gpus", 2, "--shard", "samples", "-s", 4, "-r", 160, "-o", tmp_path / "cnt.pfm", js)
    c = scenefiles.read_pfm(tmp_path / "cnt.pfm")
    assert c.shape == a.shape and abs(float(c.mean()) / float(a.mean()) - 1) < 0.1
