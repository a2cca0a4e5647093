"""Batched hair-BSDF calls through the C ABI (numpy in, numpy out).

Names follow the reference's `yocto::extension` API
(libs/yocto_extension/yocto_extension.h:115-130); each call runs the CUDA
kernels of csrc/yh_bsdf.cu — there is no host implementation here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi

_P = C.c_void_p


def _f32(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if cols is not None:
        a = a.reshape(-1, cols) if cols > 1 else a.reshape(-1)
    return a


def _ptr(a):
    return a.ctypes.data_as(_P)


def make_materials(n: int, **fields) -> C.Array:
    """Array of yh_hair_material with the reference defaults
    (yocto_extension.h:86-95) overridden per field (scalars or length-n arrays)."""
    arr = (capi.HairMaterial * n)()
    base = dict(sigma_a=(0, 0, 0), beta_m=0.3, beta_n=0.3, alpha=2.0, eta=1.55,
                color=(0, 0, 0), eumelanin=0.0, pheomelanin=0.0)
    base.update(fields)
    view = np.frombuffer(arr, dtype=np.float32).reshape(n, 12)
    cols = dict(sigma_a=slice(0, 3), beta_m=3, beta_n=4, alpha=5, eta=6, color=slice(7, 10),
                eumelanin=10, pheomelanin=11)
    for k, v in base.items():
        view[:, cols[k]] = np.asarray(v, dtype=np.float32)
    return arr


def brdf_array(n: int) -> C.Array:
    return (capi.HairBrdf * n)()


def brdf_view(brdfs) -> np.ndarray:
    """(n, 30) float32 view of an array of yh_hair_brdf."""
    return np.frombuffer(brdfs, dtype=np.float32).reshape(len(brdfs), 30)


def eval_hair_brdf(materials, v, normals, tangents):
    n = len(materials)
    v, normals, tangents = _f32(v, 1), _f32(normals, 3), _f32(tangents, 3)
    out = brdf_array(n)
    capi.check(capi.load().yh_hair_brdf_batch(C.addressof(materials), _ptr(v), _ptr(normals),
                                              _ptr(tangents), n, C.addressof(out)))
    return out


def eval_hair_scattering(brdfs, outgoing, incoming) -> np.ndarray:
    n = len(brdfs)
    wo, wi = _f32(outgoing, 3), _f32(incoming, 3)
    out = np.empty((n, 3), np.float32)
    capi.check(capi.load().yh_hair_eval_batch(C.addressof(brdfs), _ptr(wo), _ptr(wi), n, _ptr(out)))
    return out


def sample_hair_scattering(brdfs, outgoing, rn) -> np.ndarray:
    n = len(brdfs)
    wo, rn = _f32(outgoing, 3), _f32(rn, 2)
    out = np.empty((n, 3), np.float32)
    capi.check(capi.load().yh_hair_sample_batch(C.addressof(brdfs), _ptr(wo), _ptr(rn), n, _ptr(out)))
    return out


def sample_hair_scattering_pdf(brdfs, outgoing, incoming) -> np.ndarray:
    n = len(brdfs)
    wo, wi = _f32(outgoing, 3), _f32(incoming, 3)
    out = np.empty((n,), np.float32)
    capi.check(capi.load().yh_hair_pdf_batch(C.addressof(brdfs), _ptr(wo), _ptr(wi), n, _ptr(out)))
    return out


SELFTESTS = ("white_furnace_test", "white_furnace_sampled_test", "sampling_weights_test",
             "sampling_consistency_test")


def hair_selftest(which: int):
    """Runs one of the reference's four statistical tests
    (yocto_extension.cpp:555-693) on the device BSDF -> (passed, worst)."""
    passed, worst = C.c_int(0), C.c_float(0)
    capi.check(capi.load().yh_hair_selftest(which, C.byref(passed), C.byref(worst)))
    return bool(passed.value), worst.value


def rng_floats(seed: int, seq: int, n: int) -> np.ndarray:
    out = np.empty((n,), np.float32)
    capi.check(capi.load().yh_rng_floats(seed, seq, n, _ptr(out)))
    return out


def libm(fn: int, a, b=None) -> np.ndarray:
    a = _f32(a)
    b = _f32(b) if b is not None else a
    out = np.empty_like(a)
    capi.check(capi.load().yh_libm_batch(fn, _ptr(a), _ptr(b), a.size, _ptr(out)))
    return out
