"""TEST INFRASTRUCTURE — ctypes wrappers around the CPU checkers.

  RefOracle   oracle/_ref/libyh_ref.so   = the UNMODIFIED reference C++ compiled
              from /root/reference by oracle/Makefile (kind "reference").
  PortOracle  oracle/_build/libyh_oracle.so = the plain-C restatement
              oracle/yh_oracle.c (kind "port").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs import this module.  The product never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent / "yocto-hair_b200"))
from yocto_hair_b200 import _capi as capi  # noqa: E402  (struct layouts only)

REF_LIB = HERE / "_ref" / "libyh_ref.so"
PORT_LIB = HERE / "_build" / "libyh_oracle.so"
_P = C.c_void_p


def build_port() -> Path:
    subprocess.run(["make", "-s", "-C", str(HERE), "port"], check=True)
    return PORT_LIB


def build_ref() -> Path:
    """Only possible where /root/reference exists (the build container)."""
    subprocess.run(["make", "-s", "-j8", "-C", str(HERE), "ref"], check=True)
    return REF_LIB


def _ptr(a):
    return a.ctypes.data_as(_P) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class _Oracle:
    """Common surface of both checkers (same exported names modulo prefix)."""

    prefix = ""
    kind = ""

    def __init__(self, path: Path):
        if not path.exists():
            raise FileNotFoundError(path)
        self.lib = C.CDLL(str(path))
        self.path = path
        L, p = self.lib, self.prefix
        self._fn = lambda name: getattr(L, p + name)
        for name, res, args in [
            ("hair_brdf_batch", None, [_P, _P, _P, _P, C.c_int64, _P]),
            ("hair_eval_batch", None, [_P, _P, _P, C.c_int64, _P]),
            ("hair_sample_batch", None, [_P, _P, _P, C.c_int64, _P]),
            ("hair_pdf_batch", None, [_P, _P, _P, C.c_int64, _P]),
            ("hair_selftest", C.c_int, [C.c_int]),
            ("rng_floats", None, [C.c_uint64, C.c_uint64, C.c_int, _P]),
            ("libm_batch", None, [C.c_int, _P, _P, C.c_int64, _P]),
            ("scene_create", _P, [C.POINTER(capi.SceneDesc)]),
            ("scene_destroy", None, [_P]),
            ("intersect_batch", None, [_P, _P, C.c_int64, C.c_int, _P, C.c_int]),
            ("state_create", _P, [_P, C.POINTER(capi.CameraDesc), C.POINTER(capi.TraceParams)]),
            ("state_destroy", None, [_P]),
            ("state_size", None, [_P, capi.c_int_p, capi.c_int_p]),
            ("trace_samples", C.c_double, [_P, C.c_int]),
            ("state_download", None, [_P, _P, _P, _P, _P]),
            ("hardware_concurrency", C.c_int, []),
            ("shape_order", C.c_int, [_P, C.c_int, _P]),
        ]:
            f = self._fn(name)
            f.restype, f.argtypes = res, args
        if self.kind == "port":  # the library's own counter mode, restated in the port only
            f = self._fn("state_set_sample_shard")
            f.restype, f.argtypes = None, [_P, C.c_int, C.c_int]

    # ---- hair BSDF ----------------------------------------------------------
    def hair_brdf(self, materials, v, normals, tangents):
        n = len(v)
        out = (capi.HairBrdf * n)()
        self._fn("hair_brdf_batch")(C.addressof(materials), _ptr(_f32(v)), _ptr(_f32(normals)),
                                    _ptr(_f32(tangents)), n, C.addressof(out))
        return out

    def hair_eval(self, brdfs, wo, wi):
        n = len(brdfs)
        out = np.zeros((n, 3), np.float32)
        self._fn("hair_eval_batch")(C.addressof(brdfs), _ptr(_f32(wo)), _ptr(_f32(wi)), n, _ptr(out))
        return out

    def hair_sample(self, brdfs, wo, rn):
        n = len(brdfs)
        out = np.zeros((n, 3), np.float32)
        self._fn("hair_sample_batch")(C.addressof(brdfs), _ptr(_f32(wo)), _ptr(_f32(rn)), n, _ptr(out))
        return out

    def hair_pdf(self, brdfs, wo, wi):
        n = len(brdfs)
        out = np.zeros((n,), np.float32)
        self._fn("hair_pdf_batch")(C.addressof(brdfs), _ptr(_f32(wo)), _ptr(_f32(wi)), n, _ptr(out))
        return out

    def hair_selftest(self, which: int) -> bool:
        return self._fn("hair_selftest")(which) == 0

    def rng_floats(self, seed, seq, n):
        out = np.zeros((n,), np.float32)
        self._fn("rng_floats")(seed, seq, n, _ptr(out))
        return out

    def libm(self, fn: int, a, b=None):
        a = _f32(a)
        b = _f32(b) if b is not None else a
        out = np.zeros_like(a)
        self._fn("libm_batch")(fn, _ptr(a), _ptr(b), a.size, _ptr(out))
        return out

    def hardware_concurrency(self) -> int:
        return self._fn("hardware_concurrency")()

    # ---- scene / intersect / render -------------------------------------------
    def scene(self, scene):
        return _OracleScene(self, scene)


class _OracleScene:
    def __init__(self, oracle: _Oracle, scene):
        self.o = oracle
        desc, self._keep = scene.to_desc()
        self.handle = oracle._fn("scene_create")(C.byref(desc))
        if not self.handle:
            raise RuntimeError("oracle scene_create failed")

    def close(self):
        if self.handle:
            self.o._fn("scene_destroy")(self.handle)
            self.handle = None

    def intersect(self, rays: np.ndarray, find_any=False, threads=1) -> np.ndarray:
        """rays: structured/float32 (n, 8) = o, d, tmin, tmax -> hits record array."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hits = np.zeros((len(rays),), dtype=HIT_DTYPE)
        self.o._fn("intersect_batch")(self.handle, _ptr(rays), len(rays), int(find_any), _ptr(hits), threads)
        return hits

    def shape_order(self, shape: int) -> np.ndarray:
        """Primitive order of the checker's BVH over `shape` (shape->bvh->primitives after build_bvh)."""
        n = self.o._fn("shape_order")(self.handle, shape, None)
        order = np.zeros(max(n, 0), np.int32)
        self.o._fn("shape_order")(self.handle, shape, _ptr(order))
        return order

    def state(self, camera, params):
        return _OracleState(self, camera, params)


class _OracleState:
    def __init__(self, sc: _OracleScene, camera, params):
        self.sc, self.o = sc, sc.o
        cam = camera.to_desc()
        self.handle = self.o._fn("state_create")(sc.handle, C.byref(cam), C.byref(params))
        w, h = C.c_int(), C.c_int()
        self.o._fn("state_size")(self.handle, C.byref(w), C.byref(h))
        self.width, self.height = w.value, h.value

    def set_sample_shard(self, rank: int, world: int):
        """Counter-mode streams (port only; not a reference feature)."""
        self.o._fn("state_set_sample_shard")(self.handle, rank, world)
        return self

    def trace_samples(self, n: int) -> float:
        """Runs n passes of trace_samples; returns seconds inside the loop."""
        return self.o._fn("trace_samples")(self.handle, n)

    def download(self):
        n = self.width * self.height
        render = np.zeros((self.height, self.width, 4), np.float32)
        accum = np.zeros((self.height, self.width, 4), np.float32)
        samples = np.zeros((self.height, self.width), np.int32)
        rng = np.zeros((n, 2), np.uint64)
        self.o._fn("state_download")(self.handle, _ptr(render), _ptr(accum), _ptr(samples), _ptr(rng))
        return render, accum, samples, rng

    def close(self):
        if self.handle:
            self.o._fn("state_destroy")(self.handle)
            self.handle = None


HIT_DTYPE = np.dtype([("object", np.int32), ("element", np.int32), ("uv", np.float32, (2,)),
                      ("distance", np.float32), ("hit", np.int32)])


class RefOracle(_Oracle):
    prefix, kind = "ref_", "reference"

    def __init__(self):
        super().__init__(REF_LIB)


class PortOracle(_Oracle):
    prefix, kind = "yho_", "port"

    def __init__(self):
        if not PORT_LIB.exists():
            build_port()
        super().__init__(PORT_LIB)


def best_oracle() -> _Oracle:
    """The reference build when it travelled with the snapshot, else the port (YH_ORACLE=port forces the port)."""
    import os
    if os.environ.get("YH_ORACLE") == "port":
        return PortOracle()
    return RefOracle() if REF_LIB.exists() else PortOracle()
