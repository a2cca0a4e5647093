"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/yocto_hair_b200.h declares, the ctypes mirror matches the header's
struct layouts, and — without a CUDA device — every compute entry point fails
loudly with YH_ERR_NO_DEVICE (there is no CPU fallback in the product)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from yocto_hair_b200 import _capi as capi

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "yocto_hair_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(yh_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in capi.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(capi.PROTOTYPES) == names  # and nothing undeclared is bound
    assert lib.yh_abi_version() == 2


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof/offsetof printed by a C program compiled against the header == ctypes."""
    checks = {
        "yh_shape_desc": capi.ShapeDesc, "yh_texture_desc": capi.TextureDesc, "yh_material_desc": capi.MaterialDesc,
        "yh_object_desc": capi.ObjectDesc, "yh_environment_desc": capi.EnvironmentDesc,
        "yh_camera_desc": capi.CameraDesc, "yh_scene_desc": capi.SceneDesc, "yh_trace_params": capi.TraceParams,
        "yh_bvh_params": capi.BvhParams, "yh_bvh_stats": capi.BvhStats, "yh_trace_stats": capi.TraceStats,
        "yh_ray": capi.Ray, "yh_hit": capi.Hit, "yh_hair_material": capi.HairMaterial, "yh_hair_brdf": capi.HairBrdf,
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, cls in checks.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)  # the header is plain C
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in checks.items():
        assert int(out[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
    assert C.sizeof(capi.HairBrdf) == 120 and C.sizeof(capi.Hit) == 24 and C.sizeof(capi.Ray) == 32


def test_defaults_match_the_reference(lib):
    """trace_params / material defaults, yocto_pathtrace.h:188-197, :293-316."""
    p = capi.TraceParams()
    assert lib.yh_default_params(C.byref(p)) == capi.YH_OK
    assert (p.resolution, p.shader, p.samples, p.bounces, p.clamp, p.seed, p.pratio) == (720, 1, 512, 8, 100.0, 961748941, 8)
    m = capi.MaterialDesc()
    assert lib.yh_default_material(C.byref(m)) == capi.YH_OK
    assert (m.beta_m, m.beta_n, m.alpha, m.eta) == pytest.approx((0.3, 0.3, 2.0, 1.55))
    assert m.opacity == 1 and m.ior == 1.5 and m.emission_tex == -1 and m.normal_tex == -1
    assert lib.yh_default_params(None) == capi.YH_ERR_INVALID and lib.yh_last_error()


def test_shard_map_tiles_the_image(lib):
    """yh_shard_pixels: 32x32 tiles round-robin over ranks; the shards are disjoint,
    cover the image for any world size, and world = 1 owns everything."""
    for (w, h) in [(1280, 1280), (720, 405), (33, 65), (1, 1)]:
        for world in (1, 2, 3, 8):
            seen = np.zeros(w * h, np.int32)
            sizes = []
            for rank in range(world):
                n = C.c_int64()
                assert lib.yh_shard_pixels(w, h, rank, world, None, C.byref(n)) == capi.YH_OK
                px = np.zeros(n.value, np.int32)
                assert lib.yh_shard_pixels(w, h, rank, world, px.ctypes.data_as(capi.c_int_p), C.byref(n)) == capi.YH_OK
                assert len(np.unique(px)) == len(px)
                seen[px] += 1
                sizes.append(len(px))
                tiles = (px % w) // 32 + ((px // w) // 32) * ((w + 31) // 32)
                assert np.all(tiles % world == rank)
            assert np.all(seen == 1)
            if w * h >= 1280 * 1280:  # load balance at the bench resolution: within 3 %
                assert max(sizes) / min(sizes) < 1.03
    n = C.c_int64()
    assert lib.yh_shard_pixels(16, 16, 2, 2, None, C.byref(n)) == capi.YH_ERR_INVALID


def test_no_device_fails_loudly(lib):
    """Without a GPU the product must refuse, never fall back to a CPU path."""
    n = C.c_int(-1)
    assert lib.yh_device_count(C.byref(n)) == capi.YH_OK
    if n.value > 0:
        pytest.skip("a CUDA device is present")
    from yocto_hair_b200 import hair, synth
    from yocto_hair_b200.pathtrace import DeviceScene
    with pytest.raises(capi.YhError) as e:
        hair.rng_floats(1, 1, 4)
    assert e.value.status == capi.YH_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    mats = hair.make_materials(2)
    with pytest.raises(capi.YhError) as e:
        hair.eval_hair_brdf(mats, [0.1, 0.2], [[0, 0, 1]] * 2, [[1, 0, 0]] * 2)
    assert e.value.status == capi.YH_ERR_NO_DEVICE
    sc, _ = synth.curly_hair_scene(800)
    with pytest.raises(capi.YhError) as e:
        DeviceScene(sc)
    assert e.value.status == capi.YH_ERR_NO_DEVICE
    assert lib.yh_set_device(0) == capi.YH_ERR_NO_DEVICE


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", tmp_path / "libyocto_hair_b200.so")
    with pytest.raises(FileNotFoundError, match="no CPU fallback"):
        capi.load()
