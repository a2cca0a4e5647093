"""Host-side mirror of the reference's `yocto::pathtrace` entry points
(libs/yocto_pathtrace/yocto_pathtrace.h:206-230, :452-455) over the C ABI.

    scene = DeviceScene(scene_description)      # yh_scene_create
    scene.init_bvh(); scene.init_lights()       # init_bvh / init_lights
    state = scene.init_state(camera, params)    # init_state
    state.trace_samples(n)                      # n x trace_samples
    render = state.render()                     # state->render

Everything runs in the CUDA library; a missing library or device raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from .scene import Camera, Scene, make_params

_P = C.c_void_p

RAY_DTYPE = np.dtype([("o", np.float32, (3,)), ("d", np.float32, (3,)), ("tmin", np.float32), ("tmax", np.float32)])
HIT_DTYPE = np.dtype([("object", np.int32), ("element", np.int32), ("uv", np.float32, (2,)),
                      ("distance", np.float32), ("hit", np.int32)])
RAY_EPS = np.float32(1e-4)
FLT_MAX = np.float32(3.402823466e+38)


def make_rays(origins, directions, tmin=RAY_EPS, tmax=FLT_MAX) -> np.ndarray:
    o = np.asarray(origins, np.float32).reshape(-1, 3)
    d = np.asarray(directions, np.float32).reshape(-1, 3)
    rays = np.zeros((len(o),), RAY_DTYPE)
    rays["o"], rays["d"], rays["tmin"], rays["tmax"] = o, d, tmin, tmax
    return rays


class DeviceScene:
    def __init__(self, scene: Scene, device: int | None = None):
        self.lib = capi.load()
        if device is not None:
            capi.check(self.lib.yh_set_device(device))
        desc, keep = scene.to_desc()
        self.handle = _P()
        capi.check(self.lib.yh_scene_create(C.byref(desc), C.byref(self.handle)))
        del keep  # the library copied every array

    def close(self):
        if self.handle:
            self.lib.yh_scene_destroy(self.handle)
            self.handle = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_bvh(self, max_split: int = 0, leaf_size: int = 0, quality: int = 0) -> "DeviceScene":
        p = capi.BvhParams()
        p.max_split, p.leaf_size, p.quality = max_split, leaf_size, quality
        capi.check(self.lib.yh_scene_init_bvh(self.handle, C.byref(p)))
        return self

    def init_lights(self) -> "DeviceScene":
        capi.check(self.lib.yh_scene_init_lights(self.handle))
        return self

    def reference_order(self, shape: int) -> np.ndarray:
        """Primitive order of the reference's BVH over `shape`, as the device restates it (yh_scene_reference_order)."""
        n = C.c_int64()
        capi.check(self.lib.yh_scene_reference_order(self.handle, shape, None, C.byref(n)))
        order = np.zeros(n.value, np.int32)
        capi.check(self.lib.yh_scene_reference_order(self.handle, shape, order.ctypes.data_as(capi.c_int_p), C.byref(n)))
        return order

    def bvh_stats(self) -> capi.BvhStats:
        s = capi.BvhStats()
        capi.check(self.lib.yh_scene_bvh_stats(self.handle, C.byref(s)))
        return s

    def intersect(self, rays: np.ndarray, find_any: bool = False) -> np.ndarray:
        """Batched intersect_scene_bvh; rays is a RAY_DTYPE array (or (n, 8) float32)."""
        rays = np.ascontiguousarray(rays)
        n = rays.shape[0]
        hits = np.zeros((n,), HIT_DTYPE)
        capi.check(self.lib.yh_intersect_batch(self.handle, rays.ctypes.data_as(_P), n, int(find_any),
                                               hits.ctypes.data_as(_P)))
        return hits

    def intersect_timed(self, rays: np.ndarray, repeats: int = 3):
        """Rays resident in HBM; returns (hits, best kernel ms)."""
        rays = np.ascontiguousarray(rays)
        n = rays.shape[0]
        d_rays, d_hits = _P(), _P()
        capi.check(self.lib.yh_dev_alloc(C.byref(d_rays), rays.nbytes))
        capi.check(self.lib.yh_dev_alloc(C.byref(d_hits), n * HIT_DTYPE.itemsize))
        try:
            capi.check(self.lib.yh_dev_upload(d_rays, rays.ctypes.data_as(_P), rays.nbytes))
            best = None
            for _ in range(repeats):
                ms = C.c_float()
                capi.check(self.lib.yh_intersect_batch_dev(self.handle, d_rays, n, 0, d_hits, C.byref(ms)))
                best = ms.value if best is None else min(best, ms.value)
            hits = np.zeros((n,), HIT_DTYPE)
            capi.check(self.lib.yh_dev_download(hits.ctypes.data_as(_P), d_hits, hits.nbytes))
        finally:
            self.lib.yh_dev_free(d_rays)
            self.lib.yh_dev_free(d_hits)
        return hits, best

    def init_state(self, camera: Camera, params: capi.TraceParams | None = None) -> "DeviceState":
        return DeviceState(self, camera, params or make_params())


class DeviceState:
    def __init__(self, scene: DeviceScene, camera: Camera, params: capi.TraceParams):
        self.scene, self.lib = scene, scene.lib
        self.handle = _P()
        cam = camera.to_desc()
        capi.check(self.lib.yh_state_create(scene.handle, C.byref(cam), C.byref(params), C.byref(self.handle)))
        w, h = C.c_int(), C.c_int()
        capi.check(self.lib.yh_state_size(self.handle, C.byref(w), C.byref(h)))
        self.width, self.height = w.value, h.value

    def close(self):
        if self.handle:
            self.lib.yh_state_destroy(self.handle)
            self.handle = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_shard(self, rank: int, world: int) -> "DeviceState":
        capi.check(self.lib.yh_state_set_shard(self.handle, rank, world))
        return self

    def set_sample_shard(self, rank: int, world: int) -> "DeviceState":
        """Counter-mode streams + sample-range sharding (yh_state_set_sample_shard)."""
        capi.check(self.lib.yh_state_set_sample_shard(self.handle, rank, world))
        return self

    def set_sample_slots(self, slots_per_pixel: int = 0) -> "DeviceState":
        """Counter mode: several concurrent path slots per pixel (0 = auto), yh_state_set_sample_slots."""
        capi.check(self.lib.yh_state_set_sample_slots(self.handle, slots_per_pixel))
        return self

    def set_profiling(self, enabled: bool = True) -> "DeviceState":
        capi.check(self.lib.yh_state_set_profiling(self.handle, int(enabled)))
        return self

    def trace_samples(self, num_samples: int = 1) -> capi.TraceStats:
        capi.check(self.lib.yh_trace_samples(self.handle, num_samples, None))
        return self.stats()

    def stats(self) -> capi.TraceStats:
        s = capi.TraceStats()
        capi.check(self.lib.yh_state_stats(self.handle, C.byref(s)))
        return s

    def download(self, render=True, accum=True, samples=True):
        r = np.zeros((self.height, self.width, 4), np.float32) if render else None
        a = np.zeros((self.height, self.width, 4), np.float32) if accum else None
        s = np.zeros((self.height, self.width), np.int32) if samples else None
        p = lambda x: x.ctypes.data_as(_P) if x is not None else None  # noqa: E731
        capi.check(self.lib.yh_state_download(self.handle, p(r), p(a), p(s)))
        return r, a, s

    def render(self) -> np.ndarray:
        return self.download(True, False, False)[0]

    def download_into(self, render: np.ndarray) -> None:
        """Device -> caller-provided (pinned) host buffer, H x W x 4 float32."""
        capi.check(self.lib.yh_state_download(self.handle, render.ctypes.data_as(_P), None, None))

    def rng(self) -> np.ndarray:
        out = np.zeros((self.height * self.width, 2), np.uint64)
        capi.check(self.lib.yh_state_download_rng(self.handle, out.ctypes.data_as(_P)))
        return out

    def buffers_dev(self):
        """(accum_ptr, samples_ptr) device addresses, for the NVLink reduce."""
        a, s = _P(), _P()
        capi.check(self.lib.yh_state_buffers_dev(self.handle, C.byref(a), C.byref(s)))
        return a.value, s.value


def trace_image(scene: Scene, camera: Camera, params: capi.TraceParams | None = None, max_split: int = 0):
    """Convenience with the upstream `trace_image` meaning (libs/yocto/yocto_trace.h:275):
    build, then render params.samples samples per pixel -> (H, W, 4) float32."""
    params = params or make_params()
    ds = DeviceScene(scene).init_bvh(max_split).init_lights()
    st = ds.init_state(camera, params)
    st.trace_samples(params.samples)
    img = st.render()
    st.close()
    ds.close()
    return img


class DeviceGroup:
    """Several GPUs of one node in one process (yh_group_*): replicated scene, one state per
    GPU, fused peer-memory reduce + resolve on download."""

    def __init__(self, scene: Scene, devices=None, num_devices: int | None = None, max_split: int = 0, quality: int = 0):
        self.lib = capi.load()
        desc, keep = scene.to_desc()
        bp = capi.BvhParams()
        bp.max_split, bp.quality = max_split, quality
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            n = len(devices)
        else:
            arr, n = None, int(num_devices or 1)
        self.handle = _P()
        capi.check(self.lib.yh_group_create(C.byref(desc), C.byref(bp), arr, n, C.byref(self.handle)))
        del keep
        self.num_devices = n
        self.width = self.height = 0

    def close(self):
        if self.handle:
            self.lib.yh_group_destroy(self.handle)
            self.handle = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_state(self, camera: Camera, params: capi.TraceParams | None = None, shard: str = "tiles") -> "DeviceGroup":
        cam = camera.to_desc()
        params = params or make_params()
        mode = {"tiles": capi.YH_SHARD_TILES, "samples": capi.YH_SHARD_SAMPLES}[shard]
        capi.check(self.lib.yh_group_init_state(self.handle, C.byref(cam), C.byref(params), mode))
        w, h = C.c_int(), C.c_int()
        capi.check(self.lib.yh_group_size(self.handle, C.byref(w), C.byref(h)))
        self.width, self.height = w.value, h.value
        return self

    def set_reduce(self, mode: str) -> "DeviceGroup":
        capi.check(self.lib.yh_group_set_reduce(self.handle, {"peer": capi.YH_REDUCE_PEER, "nccl": capi.YH_REDUCE_NCCL}[mode]))
        return self

    def trace_samples(self, num_samples: int = 1) -> capi.TraceStats:
        capi.check(self.lib.yh_group_trace_samples(self.handle, num_samples, None))
        return self.stats()[0]

    def stats(self):
        s, ms = capi.TraceStats(), C.c_float()
        capi.check(self.lib.yh_group_stats(self.handle, C.byref(s), C.byref(ms)))
        return s, ms.value

    def download(self, render=True, accum=True, samples=True):
        r = np.zeros((self.height, self.width, 4), np.float32) if render else None
        a = np.zeros((self.height, self.width, 4), np.float32) if accum else None
        s = np.zeros((self.height, self.width), np.int32) if samples else None
        p = lambda x: x.ctypes.data_as(_P) if x is not None else None  # noqa: E731
        capi.check(self.lib.yh_group_download(self.handle, p(r), p(a), p(s)))
        return r, a, s

    def download_into(self, render: np.ndarray) -> None:
        capi.check(self.lib.yh_group_download(self.handle, render.ctypes.data_as(_P), None, None))
