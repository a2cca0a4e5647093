"""The surface lobes, refraction and volume functions the device uses (csrc/yh_surface.cuh) compiled for the HOST
and compared, function by function and bit for bit, with the reference's own yocto_math.h functions (called through
oracle/_ref/libyh_ref.so::ref_surface_batch) on 2 x 10^5 seeded tuples each, plus hand-picked corner inputs
(grazing and back-facing directions, roughness at its clamp, ior = 1, total internal reflection).  CPU only; the
whole-material logic (eval_brdf, the delta branch, the volume stack) is pinned by the bit-identical renders of
tests/test_trace_gpu.py."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "yocto-hair_b200" / "csrc" / "libyh_surface_host.so"
REF = ROOT / "oracle" / "_ref" / "libyh_ref.so"
_P = C.c_void_p

FUNCTIONS = ["fresnel_dielectric", "fresnel_conductor", "reflectivity_to_eta", "microfacet_distribution",
             "microfacet_shadowing", "sample_microfacet", "eval_microfacet_reflection", "eval_microfacet_reflection(metal)",
             "eval_microfacet_transmission", "eval_microfacet_refraction", "sample_microfacet_reflection",
             "sample_microfacet_transmission", "sample_microfacet_refraction", "sample_microfacet_reflection_pdf",
             "sample_microfacet_transmission_pdf", "sample_microfacet_refraction_pdf", "eval_delta_refraction",
             "sample_delta_refraction", "sample_delta_refraction_pdf", "eval_phasefunction", "sample_phasefunction",
             "sample_transmittance", "sample_transmittance_pdf", "eval_transmittance", "sample_hemisphere_cos", "refract"]


@pytest.fixture(scope="module")
def libs():
    if not HOST.exists() or not REF.exists():
        pytest.skip("libyh_surface_host.so / oracle/_ref/libyh_ref.so not built (run __graft_entry__.build() where /root/reference exists)")
    h, r = C.CDLL(str(HOST)), C.CDLL(str(REF))
    h.yh_host_surface_batch.argtypes = [C.c_int, C.c_long, _P, _P]
    r.ref_surface_batch.argtypes = [C.c_int, C.c_int64, _P, _P]
    h.yh_host_surface_batch.restype = r.ref_surface_batch.restype = None
    return h, r


def unit(rng, n):
    v = rng.normal(size=(n, 3))
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def tuples(n=200_000, seed=17):
    rng = np.random.default_rng(seed)
    t = np.zeros((n, 20), np.float32)
    t[:, 0:3], t[:, 3:6], t[:, 6:9] = unit(rng, n), unit(rng, n), unit(rng, n)
    t[:, 9:11] = rng.random((n, 2))
    t[:, 11] = rng.random(n)
    t[:, 12] = rng.choice([1.0, 1.0005, 1.33, 1.5, 2.4], n)                       # ior (|ior - 1| < 1e-3 branch included)
    t[:, 13] = rng.choice([0.03 * 0.03, 0.01, 0.1, 0.25, 0.6, 1.0], n)            # roughness (already squared / clamped)
    t[:, 14:17] = rng.uniform(0.0, 3.0, (n, 3))                                   # eta / reflectivity / density
    t[:, 17:19] = rng.uniform(0.0, 4.0, (n, 2))
    t[:, 19] = rng.choice([0.0, 5e-4, -0.8, 0.3, 0.95], n)                        # anisotropy (isotropic branch included)
    # corner inputs: incoming mirrored / opposite / perpendicular, outgoing along and against the normal
    k = n // 10
    t[:k, 6:9] = -t[:k, 3:6]
    t[k:2 * k, 3:6] = t[k:2 * k, 0:3]
    t[2 * k:3 * k, 3:6] = -t[2 * k:3 * k, 0:3]
    d = np.cross(t[3 * k:4 * k, 0:3], t[3 * k:4 * k, 3:6])
    t[3 * k:4 * k, 6:9] = d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
    t[4 * k:5 * k, 9:11] = rng.choice([0.0, 1.0 - 2.0 ** -24, 0.5], (k, 2))
    return np.ascontiguousarray(t)


@pytest.mark.parametrize("fn", range(len(FUNCTIONS)), ids=FUNCTIONS)
def test_surface_function_bit_exact(fn, libs):
    host, ref = libs
    t = tuples()
    n = len(t)
    a, b = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    host.yh_host_surface_batch(fn, n, t.ctypes.data_as(_P), a.ctypes.data_as(_P))
    ref.ref_surface_batch(fn, n, t.ctypes.data_as(_P), b.ctypes.data_as(_P))
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    bad = np.nonzero(~same.all(axis=1))[0]
    assert bad.size == 0, "%s: %d of %d differ, first %s -> host %s ref %s" % (
        FUNCTIONS[fn], bad.size, n, t[bad[0]], a[bad[0]], b[bad[0]])
    assert np.count_nonzero(b) > 0  # the tuples do reach the function's non-trivial branches
