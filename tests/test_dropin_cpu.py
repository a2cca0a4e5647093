"""Drop-in plumbing on CPU: the harness writes a synthetic hair scene in the
reference's own on-disk formats (scene JSON + binary PLY polylines), the drop-in
A/B test tool host/_build/yscenetrace_ab (the drop-in command line compiled with
-DYH_AB_TOOL) loads it through the reference's unmodified yocto_sceneio and renders
it with `--backend cpu` (the reference's own functions).  The image must equal, bit for bit, the render of the same
scene handed to the reference IN MEMORY — which pins the file writers, the
io -> ptr scene conversion of the app and the flattened C-ABI description to
each other.  (The product binary yscenetrace_b200 is compared with it in
tests/test_dropin_gpu.py; it has no CPU backend, which is tested here too.)"""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle as orc
from yocto_hair_b200 import scenefiles, synth
from yocto_hair_b200.scene import make_params

ROOT = Path(__file__).resolve().parent.parent
APP = ROOT / "yocto-hair_b200" / "host" / "_build" / "yscenetrace_b200"
AB = ROOT / "yocto-hair_b200" / "host" / "_build" / "yscenetrace_ab"


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
    run = subprocess.run([str(AB), "--backend", "cpu", "-s", "2", "-r", "72", "-o", str(out), str(js)],
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
    run = subprocess.run([str(APP), "-s", "1", "-r", "32", "-o", str(tmp_path / "x.pfm"), str(js)],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode != 0  # no device here -> YH_ERR_NO_DEVICE; on a GPU box -> YH_ERR_UNSUPPORTED; never a silent CPU render
    assert not (tmp_path / "x.pfm").exists()


def test_product_binary_has_no_cpu_backend(tmp_path):
    """north_star: no CPU fallback.  The product command line does not know `--backend`; only the A/B test tool does."""
    need_app()
    scene, cam = synth.curly_hair_scene(500, sky=False)
    js = scenefiles.write_scene(tmp_path, "nocpu", scene, cam)
    run = subprocess.run([str(APP), "--backend", "cpu", "-s", "1", "-r", "32", "-o", str(tmp_path / "x.pfm"), str(js)],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode != 0 and not (tmp_path / "x.pfm").exists()
    syms = subprocess.run(["nm", "-C", str(APP)], capture_output=True, text=True).stdout
    assert "reference_backend" not in syms


# ---- scene files as first-class inputs: host/yh_sceneio (bulk PLY reader + restated
# JSON schema) must fill the reference's sceneio::model exactly as the reference's own loader does.
# host/_build/yh_loadcheck loads a scene with both and compares every field bit for bit.
CHECK = ROOT / "yocto-hair_b200" / "host" / "_build" / "yh_loadcheck"


def loadcheck(js):
    import json
    if not CHECK.exists():
        pytest.skip("host/_build/yh_loadcheck not built (needs /root/reference at build time)")
    run = subprocess.run([str(CHECK), str(js)], capture_output=True, text=True, timeout=600)
    return run, (json.loads(run.stdout.strip().splitlines()[-1]) if run.stdout.strip().startswith("{") else None)


@pytest.mark.parametrize("name,make", [
    ("curly", lambda: synth.curly_hair_scene(20_000, sky=True)),         # hair polylines + the sky texture (.hdr)
    ("straight", lambda: synth.straight_hair_scene(5_000, sky=False)),   # + two area lights (triangles)
    ("block", lambda: synth.hair_block_scene(2_000)),                    # + diffuse sphere, 16:9 camera
    ("curls", lambda: synth.hair_curls_scene(500)),                      # four objects share one shape
])
def test_fast_loader_equals_the_reference_loader(name, make, tmp_path):
    scene, cam = make()
    js = scenefiles.write_scene(tmp_path, name, scene, cam)
    run, res = loadcheck(js)
    assert run.returncode == 0 and res["equal"], run.stdout[-1000:] + run.stderr[-2000:]
    assert res["segments"] == sum(len(s.lines) for s in scene.shapes)


def _ply(path, header, body=b""):
    Path(path).write_bytes(("ply\n" + header + "end_header\n").encode() + body)


def test_fast_loader_schema_corners(tmp_path):
    """lookat cameras / environments / objects, instances, shared and scalar textures, a default material,
    quads + polygons, and PLY variants outside the bulk path (ascii, uchar colours, double positions)."""
    import json
    import struct
    d = tmp_path
    for sub in ("shapes", "textures", "instances"):
        (d / sub).mkdir()
    v4 = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 0.5, 1]], "<f4")
    # faces with a quad, a triangle and a pentagon -> quads
    faces = b"".join(bytes([len(f)]) + np.array(f, "<i4").tobytes() for f in ([0, 1, 2, 3], [0, 1, 4], [0, 1, 2, 3, 4]))
    _ply(d / "shapes" / "poly.ply", "format binary_little_endian 1.0\ncomment corner cases # trailing\nobj_info x\n"
         "element vertex 5\nproperty float x\nproperty float y\nproperty float z\nproperty float u\nproperty float v\n"
         "element face 3\nproperty list uchar int vertex_indices\n",
         np.concatenate([v4, v4[:, :2]], axis=1).astype("<f4").tobytes() + faces)
    # triangles only + points + an element the loaders ignore, extra vertex property in the middle of the row
    rows = b"".join(struct.pack("<fffBf", *p, 7, 0.25) for p in v4.tolist())
    _ply(d / "shapes" / "tris.ply", "format binary_little_endian 1.0\nelement vertex 5\nproperty float x\nproperty float y\n"
         "property float z\nproperty uchar flag\nproperty float radius\nelement face 2\nproperty list uchar int vertex_indices\n"
         "element point 1\nproperty list uchar int vertex_indices\nelement extra 2\nproperty short a\nproperty double b\n",
         rows + bytes([3]) + np.array([0, 1, 2], "<i4").tobytes() + bytes([3]) + np.array([0, 2, 3], "<i4").tobytes()
         + bytes([2]) + np.array([3, 4], "<i4").tobytes() + struct.pack("<hdhd", 1, 2.0, 3, 4.0))
    # hair without radius (-> default 0.001), s/t texcoords
    _ply(d / "shapes" / "hair.ply", "format binary_little_endian 1.0\nelement vertex 5\nproperty float x\nproperty float y\n"
         "property float z\nproperty float s\nproperty float t\nelement line 2\nproperty list uchar int vertex_indices\n",
         np.concatenate([v4, v4[:, :2]], axis=1).astype("<f4").tobytes()
         + bytes([3]) + np.array([0, 1, 2], "<i4").tobytes() + bytes([2]) + np.array([3, 4], "<i4").tobytes())
    # outside the bulk path: ascii, and uchar colours with double positions
    (d / "shapes" / "ascii.ply").write_text("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\n"
                                            "property float z\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n"
                                            "0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n")
    _ply(d / "shapes" / "typed.ply", "format binary_little_endian 1.0\nelement vertex 3\nproperty double x\nproperty double y\n"
         "property double z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nelement face 1\n"
         "property list uchar uint vertex_indices\n",
         b"".join(struct.pack("<dddBBB", *p, 10, 20, 30) for p in ([0, 0, 0], [1, 0, 0], [0, 1, 0])) + bytes([3]) + np.array([0, 1, 2], "<u4").tobytes())
    _ply(d / "instances" / "grid.ply", "format binary_little_endian 1.0\nelement instance 2\n"
         + "".join(f"property float {n}\n" for n in "xx xy xz yx yy yz zx zy zz ox oy oz".split()),
         np.array([[1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], [1, 0, 0, 0, 1, 0, 0, 0, 1, 2, 0, 0]], "<f4").tobytes())
    rng = np.random.default_rng(3)
    scenefiles.write_hdr(d / "textures" / "sky.hdr", rng.random((8, 16, 3)).astype(np.float32) + 0.1)
    scenefiles.write_hdr(d / "textures" / "both.hdr", rng.random((4, 4, 3)).astype(np.float32) + 0.1)
    js = {
        "asset": {"copyright": "test"},
        "cameras": {"b": {"lookat": [0, 1, 5, 0, 1, 0, 0, 1, 0], "lens": 0.05, "aspect": 1, "film": 0.036, "aperture": 0.01},
                    "a": {"frame": [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 3], "orthographic": True}},
        "environments": {"sky": {"emission": [1, 2, 3], "emission_tex": "sky", "lookat": [0, 0, 0, 1, 0, 0, 0, 1, 0]},
                         "flat": {"emission": [0.5, 0.5, 0.5], "frame": [0, 0, 1, 0, 1, 0, -1, 0, 0, 0, 0, 0]}},
        "materials": {"hair": {"eumelanin": 1.3, "pheomelanin": 0.2, "sigma_a": [0.1, 0.2, 0.3], "beta_m": 0.25, "beta_n": 0.4,
                               "alpha": 3, "eta": 1.5},
                      "shiny": {"color": [0.8, 0.6, 0.4], "color_tex": "both", "roughness_tex": "both", "specular": 1, "roughness": 0.3,
                                "metallic": 0.1, "ior": 1.4, "transmission": 0.2, "thin": False, "trdepth": 0.02, "opacity": 0.9,
                                "scattering": [0.1, 0.1, 0.1], "scanisotropy": -0.5, "emission_tex": "sky", "normal_tex": "both",
                                "subdivisions": 3, "smooth": False, "coat": 0.1, "translucency": 0.05, "displacement": 0.5},
                      "light": {"emission": [20, 20, 20]}},
        "objects": {"z-hair": {"shape": "hair", "material": "hair", "lookat": [0, 0, 0, 0, 0, 1, 0, 1, 0]},
                    "m-poly": {"shape": "poly", "material": "shiny", "frame": [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 2, 0], "instance": "grid"},
                    "a-tris": {"shape": "tris", "material": "light"},
                    "b-none": {"shape": "ascii"},                       # no material -> the default one is added
                    "c-typed": {"shape": "typed", "material": ""},
                    "d-again": {"shape": "hair", "material": "hair"}}}
    (d / "corner.json").write_text(json.dumps(js))
    run, res = loadcheck(d / "corner.json")
    assert run.returncode == 0 and res["equal"], run.stdout[-1000:] + run.stderr[-2000:]
    assert res["segments"] == 3
    # no cameras -> the default camera is added by both
    del js["cameras"]
    (d / "nocam.json").write_text(json.dumps(js))
    run, res = loadcheck(d / "nocam.json")
    assert run.returncode == 0 and res["equal"], run.stdout[-1000:] + run.stderr[-2000:]


def test_fast_loader_errors_are_loud(tmp_path):
    import json
    (tmp_path / "shapes").mkdir()
    base = {"objects": {"o": {"shape": "missing"}}}
    (tmp_path / "s.json").write_text(json.dumps(base))
    run, _ = loadcheck(tmp_path / "s.json")
    assert run.returncode != 0 and "missing.ply" in run.stderr
    # truncated body
    _ply(tmp_path / "shapes" / "cut.ply", "format binary_little_endian 1.0\nelement vertex 4\nproperty float x\nproperty float y\n"
         "property float z\n", np.zeros(5, "<f4").tobytes())
    (tmp_path / "c.json").write_text(json.dumps({"objects": {"o": {"shape": "cut"}}}))
    run, _ = loadcheck(tmp_path / "c.json")
    assert run.returncode != 0 and "read error" in run.stderr
    (tmp_path / "m.json").write_text(json.dumps({"objects": {"o": {"shape": "cut", "material": "nope"}}}))
    run, _ = loadcheck(tmp_path / "m.json")
    assert run.returncode != 0 and "missing material nope" in run.stderr
    (tmp_path / "bad.json").write_text("{ not json")
    run, _ = loadcheck(tmp_path / "bad.json")
    assert run.returncode != 0 and "parse error" in run.stderr
