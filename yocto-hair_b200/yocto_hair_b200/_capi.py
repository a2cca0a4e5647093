"""ctypes view of include/yocto_hair_b200.h (the C ABI of the CUDA library).

This is plumbing for the Python test/bench harness only; the product is the
shared library `csrc/libyocto_hair_b200.so`.  There is no fallback: if the
library is missing, `load()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)

YH_OK, YH_ERR_INVALID, YH_ERR_UNSUPPORTED, YH_ERR_NO_DEVICE, YH_ERR_CUDA, YH_ERR_STATE = range(6)
YH_SHARD_TILES, YH_SHARD_SAMPLES = 0, 1
YH_REDUCE_PEER, YH_REDUCE_NCCL = 0, 1


class ShapeDesc(C.Structure):
    _fields_ = [
        ("positions", c_float_p), ("normals", c_float_p), ("texcoords", c_float_p),
        ("radius", c_float_p), ("lines", c_int_p), ("triangles", c_int_p),
        ("num_positions", C.c_int), ("num_normals", C.c_int), ("num_texcoords", C.c_int),
        ("num_radius", C.c_int), ("num_lines", C.c_int), ("num_triangles", C.c_int),
        ("num_points", C.c_int), ("points", c_int_p),
    ]


class TextureDesc(C.Structure):
    _fields_ = [("colorf", c_float_p), ("width", C.c_int), ("height", C.c_int)]


class MaterialDesc(C.Structure):
    _fields_ = [
        ("emission", C.c_float * 3), ("color", C.c_float * 3),
        ("specular", C.c_float), ("roughness", C.c_float), ("metallic", C.c_float), ("ior", C.c_float),
        ("spectint", C.c_float * 3), ("transmission", C.c_float),
        ("scattering", C.c_float * 3), ("scanisotropy", C.c_float), ("trdepth", C.c_float),
        ("opacity", C.c_float), ("thin", C.c_int),
        ("eumelanin", C.c_float), ("pheomelanin", C.c_float), ("sigma_a", C.c_float * 3),
        ("beta_m", C.c_float), ("beta_n", C.c_float), ("alpha", C.c_float), ("eta", C.c_float),
        ("emission_tex", C.c_int), ("color_tex", C.c_int), ("specular_tex", C.c_int),
        ("metallic_tex", C.c_int), ("roughness_tex", C.c_int), ("transmission_tex", C.c_int),
        ("spectint_tex", C.c_int), ("scattering_tex", C.c_int), ("opacity_tex", C.c_int),
        ("normal_tex", C.c_int),
    ]


class ObjectDesc(C.Structure):
    _fields_ = [("frame", C.c_float * 12), ("shape", C.c_int), ("material", C.c_int)]


class EnvironmentDesc(C.Structure):
    _fields_ = [("frame", C.c_float * 12), ("emission", C.c_float * 3), ("emission_tex", C.c_int)]


class CameraDesc(C.Structure):
    _fields_ = [("frame", C.c_float * 12), ("lens", C.c_float), ("film", C.c_float * 2),
                ("focus", C.c_float), ("aperture", C.c_float)]


class SceneDesc(C.Structure):
    _fields_ = [
        ("shapes", C.POINTER(ShapeDesc)), ("textures", C.POINTER(TextureDesc)),
        ("materials", C.POINTER(MaterialDesc)), ("objects", C.POINTER(ObjectDesc)),
        ("environments", C.POINTER(EnvironmentDesc)),
        ("num_shapes", C.c_int), ("num_textures", C.c_int), ("num_materials", C.c_int),
        ("num_objects", C.c_int), ("num_environments", C.c_int),
    ]


class TraceParams(C.Structure):
    _fields_ = [("resolution", C.c_int), ("shader", C.c_int), ("samples", C.c_int),
                ("bounces", C.c_int), ("clamp", C.c_float), ("seed", C.c_uint64),
                ("noparallel", C.c_int), ("pratio", C.c_int)]


class BvhParams(C.Structure):
    _fields_ = [("max_split", C.c_int), ("leaf_size", C.c_int), ("quality", C.c_int), ("reserved", C.c_int * 5)]


class BvhStats(C.Structure):
    _fields_ = [("num_primitives", C.c_int64), ("num_references", C.c_int64),
                ("num_nodes", C.c_int64), ("node_bytes", C.c_int64), ("prim_bytes", C.c_int64)]


class TraceStats(C.Structure):
    _fields_ = [("samples", C.c_int64), ("rays", C.c_int64), ("light_rays", C.c_int64),
                ("iterations", C.c_int64), ("launches", C.c_int64),
                ("device_ms", C.c_float), ("traverse_ms", C.c_float), ("shade_ms", C.c_float),
                ("nodes_visited", C.c_int64), ("prims_tested", C.c_int64)]


class Ray(C.Structure):
    _fields_ = [("o", C.c_float * 3), ("d", C.c_float * 3), ("tmin", C.c_float), ("tmax", C.c_float)]


class Hit(C.Structure):
    _fields_ = [("object", C.c_int), ("element", C.c_int), ("uv", C.c_float * 2),
                ("distance", C.c_float), ("hit", C.c_int)]


class HairMaterial(C.Structure):
    _fields_ = [("sigma_a", C.c_float * 3), ("beta_m", C.c_float), ("beta_n", C.c_float),
                ("alpha", C.c_float), ("eta", C.c_float), ("color", C.c_float * 3),
                ("eumelanin", C.c_float), ("pheomelanin", C.c_float)]


class HairBrdf(C.Structure):
    _fields_ = [("sigma_a", C.c_float * 3), ("alpha", C.c_float), ("eta", C.c_float), ("h", C.c_float),
                ("v", C.c_float * 4), ("s", C.c_float), ("sin_2k_alpha", C.c_float * 3),
                ("cos_2k_alpha", C.c_float * 3), ("gamma_o", C.c_float),
                ("world_to_brdf", C.c_float * 12)]


# symbol -> (restype, argtypes); every symbol include/yocto_hair_b200.h declares
_P = C.c_void_p
PROTOTYPES = {
    "yh_abi_version": (C.c_int, []),
    "yh_last_error": (C.c_char_p, []),
    "yh_device_count": (C.c_int, [c_int_p]),
    "yh_set_device": (C.c_int, [C.c_int]),
    "yh_default_params": (C.c_int, [C.POINTER(TraceParams)]),
    "yh_default_material": (C.c_int, [C.POINTER(MaterialDesc)]),
    "yh_scene_create": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(_P)]),
    "yh_scene_destroy": (None, [_P]),
    "yh_scene_init_bvh": (C.c_int, [_P, C.POINTER(BvhParams)]),
    "yh_scene_init_lights": (C.c_int, [_P]),
    "yh_scene_reference_order": (C.c_int, [_P, C.c_int, c_int_p, C.POINTER(C.c_int64)]),
    "yh_scene_bvh_stats": (C.c_int, [_P, C.POINTER(BvhStats)]),
    "yh_intersect_batch": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P]),
    "yh_intersect_batch_dev": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, c_float_p]),
    "yh_state_create": (C.c_int, [_P, C.POINTER(CameraDesc), C.POINTER(TraceParams), C.POINTER(_P)]),
    "yh_state_destroy": (None, [_P]),
    "yh_state_size": (C.c_int, [_P, c_int_p, c_int_p]),
    "yh_state_set_shard": (C.c_int, [_P, C.c_int, C.c_int]),
    "yh_state_set_sample_shard": (C.c_int, [_P, C.c_int, C.c_int]),
    "yh_state_set_sample_slots": (C.c_int, [_P, C.c_int]),
    "yh_shard_pixels": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_int_p, C.POINTER(C.c_int64)]),
    "yh_trace_samples": (C.c_int, [_P, C.c_int, _P]),
    "yh_state_stats": (C.c_int, [_P, C.POINTER(TraceStats)]),
    "yh_state_set_profiling": (C.c_int, [_P, C.c_int]),
    "yh_state_download": (C.c_int, [_P, _P, _P, _P]),
    "yh_state_buffers_dev": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "yh_resolve_dev": (C.c_int, [_P, _P, C.c_int64, _P]),
    "yh_state_download_rng": (C.c_int, [_P, _P]),
    "yh_group_create": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(BvhParams), c_int_p, C.c_int, C.POINTER(_P)]),
    "yh_group_destroy": (None, [_P]),
    "yh_group_num_devices": (C.c_int, [_P, c_int_p]),
    "yh_group_scene": (C.c_int, [_P, C.c_int, C.POINTER(_P), c_int_p]),
    "yh_group_init_state": (C.c_int, [_P, C.POINTER(CameraDesc), C.POINTER(TraceParams), C.c_int]),
    "yh_group_set_reduce": (C.c_int, [_P, C.c_int]),
    "yh_group_size": (C.c_int, [_P, c_int_p, c_int_p]),
    "yh_group_trace_samples": (C.c_int, [_P, C.c_int, _P]),
    "yh_group_stats": (C.c_int, [_P, C.POINTER(TraceStats), c_float_p]),
    "yh_group_download": (C.c_int, [_P, _P, _P, _P]),
    "yh_hair_brdf_batch": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P]),
    "yh_hair_eval_batch": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "yh_hair_sample_batch": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "yh_hair_pdf_batch": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "yh_hair_fused_dev": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P, _P, _P, c_float_p]),
    "yh_hair_selftest": (C.c_int, [C.c_int, c_int_p, c_float_p]),
    "yh_rng_floats": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int, _P]),
    "yh_libm_batch": (C.c_int, [C.c_int, _P, _P, C.c_int64, _P]),
    "yh_dev_alloc": (C.c_int, [C.POINTER(_P), C.c_int64]),
    "yh_dev_free": (C.c_int, [_P]),
    "yh_dev_upload": (C.c_int, [_P, _P, C.c_int64]),
    "yh_dev_download": (C.c_int, [_P, _P, C.c_int64]),
    "yh_dev_synchronize": (C.c_int, []),
}

PACKAGE_DIR = Path(__file__).resolve().parent.parent
LIB_PATH = Path(os.environ.get("YH_B200_LIB", PACKAGE_DIR / "csrc" / "libyocto_hair_b200.so"))

_lib = None


class YhError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"yocto_hair_b200 status {status}: {message}")
        self.status = status


def load() -> C.CDLL:
    """Loads the CUDA library; raises if it was not built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                "(yocto-hair_b200/csrc/Makefile); there is no CPU fallback")
        lib = C.CDLL(str(LIB_PATH))
        lenient = os.environ.get("YH_B200_LIB_LENIENT") == "1"  # experiments with libraries built from older sources
        for name, (res, args) in PROTOTYPES.items():
            if lenient and not hasattr(lib, name):
                continue
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(status: int) -> None:
    if status != YH_OK:
        raise YhError(status, (load().yh_last_error() or b"").decode())
