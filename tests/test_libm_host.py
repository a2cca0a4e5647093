"""The device libm restatement (csrc/yh_libm.cuh: glibc's expf / logf / sinf / cosf /
atan2f / acosf algorithms, no FMA contraction) compiled for the HOST and compared
bit for bit with the glibc the reference oracle links.  CPU only."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

LIB = Path(__file__).resolve().parent.parent / "yocto-hair_b200" / "csrc" / "libyh_libm_host.so"
_P = C.c_void_p


@pytest.fixture(scope="module")
def hostlib():
    if not LIB.exists():
        pytest.skip("libyh_libm_host.so not built (run __graft_entry__.build())")
    lib = C.CDLL(str(LIB))
    lib.yh_host_libm_batch.argtypes = [C.c_int, C.c_int, _P, _P, C.c_long, _P]
    lib.yh_host_libm_batch.restype = None
    return lib


def run(lib, fn, which, a, b=None):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b if b is not None else a, np.float32)
    out = np.empty_like(a)
    lib.yh_host_libm_batch(fn, which, a.ctypes.data_as(_P), b.ctypes.data_as(_P), a.size, out.ctypes.data_as(_P))
    return out


def specials():
    return np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 1e-30, -1e-30, 1e-45, 3.4e38, -3.4e38, np.inf, -np.inf, np.nan,
                     np.pi, -np.pi, np.pi / 2, 2 * np.pi, 88.7, -103.9, 0.70710678, 1 - 1e-7], np.float32)


@pytest.mark.parametrize("fn,name,lo,hi", [
    (0, "expf", -104.0, 89.0), (1, "logf", 1e-38, 3e38), (2, "sinf", -200.0, 200.0), (3, "cosf", -200.0, 200.0),
    (5, "acosf", -1.0, 1.0)])
def test_unary_bit_exact(hostlib, fn, name, lo, hi):
    rng = np.random.default_rng(fn)
    n = 1_000_000
    x = rng.uniform(lo, hi, n) if name != "logf" else np.exp(rng.uniform(-87, 88, n))
    x = np.concatenate([x.astype(np.float32), specials(), rng.uniform(-1, 1, 100_000).astype(np.float32) * np.float32(1e-3)])
    ours, libc = run(hostlib, fn, 0, x), run(hostlib, fn, 1, x)
    both_nan = np.isnan(ours) & np.isnan(libc)
    bad = (ours.view(np.uint32) != libc.view(np.uint32)) & ~both_nan
    assert not bad.any(), f"{name}: {bad.sum()} mismatches, first x={x[bad][:4]} ours={ours[bad][:4]} libc={libc[bad][:4]}"


def test_large_arguments_reduce_like_glibc(hostlib):
    x = np.float32(10.0) ** np.random.default_rng(9).uniform(2, 9, 200_000).astype(np.float32)
    for fn in (2, 3):
        assert np.array_equal(run(hostlib, fn, 0, x).view(np.uint32), run(hostlib, fn, 1, x).view(np.uint32))


def test_atan2f_bit_exact(hostlib):
    rng = np.random.default_rng(4)
    n = 1_000_000
    a = np.concatenate([rng.normal(size=n).astype(np.float32), np.repeat(specials(), len(specials()))])
    b = np.concatenate([(rng.normal(size=n) * np.exp(rng.uniform(-20, 20, n))).astype(np.float32),
                        np.tile(specials(), len(specials()))])
    ours, libc = run(hostlib, 4, 0, a, b), run(hostlib, 4, 1, a, b)
    both_nan = np.isnan(ours) & np.isnan(libc)
    bad = (ours.view(np.uint32) != libc.view(np.uint32)) & ~both_nan
    assert not bad.any(), f"atan2f: {bad.sum()} mismatches at {a[bad][:4]}, {b[bad][:4]}"


def test_double_sinh_is_glibcs_over_the_domain_of_mp(hostlib):
    """mp() (yocto_extension.cpp:205-206) divides by sinh(1 / v) * 2 v in DOUBLE.  yh_libm.cuh restates glibc's
    e_sinh.c over s_expm1.c (the FMA variant its ifunc selects on the oracle's CPU); here the DOUBLE results are compared
    bit for bit with glibc's sinh for a dense sample of the float arguments v in (0.1, 110] - every 64th float, 1.3 M
    values; the GPU test walks all 85 M - and outside it."""
    hostlib.yh_host_sinh_batch.argtypes = [C.c_int, _P, C.c_long, _P]
    hostlib.yh_host_sinh_batch.restype = None

    def sinh_bits(which, v):
        v = np.ascontiguousarray(v, np.float32)
        out = np.zeros(len(v), np.uint64)
        hostlib.yh_host_sinh_batch(which, v.ctypes.data_as(_P), len(v), out.ctypes.data_as(_P))
        return out

    b0, b1 = np.float32(0.1).view(np.uint32), np.float32(110.0).view(np.uint32)
    v = np.arange(int(b0), int(b1) + 1, 64, dtype=np.uint32).view(np.float32)
    assert np.array_equal(sinh_bits(0, v), sinh_bits(1, v))
    rng = np.random.default_rng(4)
    wide = np.concatenate([rng.uniform(0.046, 0.1, 200_000), rng.uniform(110, 1e6, 200_000), np.exp(rng.uniform(-3, 30, 200_000))]).astype(np.float32)
    assert np.array_equal(sinh_bits(0, wide), sinh_bits(1, wide))
