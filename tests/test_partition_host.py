"""csrc/yh_partition.h: the parallel replay of std::partition must produce libstdc++'s own permutation (the
reference's BVH leaf order, and with it the winner of exact-distance ties, depends on it).  CPU only."""
import ctypes as C
from pathlib import Path

import pytest

LIB = Path(__file__).resolve().parent.parent / "yocto-hair_b200" / "csrc" / "libyh_partition_host.so"


@pytest.fixture(scope="module")
def lib():
    if not LIB.exists():
        pytest.skip("libyh_partition_host.so not built (run __graft_entry__.build())")
    h = C.CDLL(str(LIB))
    h.yh_host_partition_check.argtypes = [C.c_long, C.c_ulonglong, C.c_int, C.c_int, C.c_float, C.c_int]
    h.yh_host_partition_check.restype = C.c_long
    return h


@pytest.mark.parametrize("threads", [2, 3, 8])
@pytest.mark.parametrize("n", [8_192, 100_003, 1_000_000])
def test_partition_replay_equals_std_partition(lib, n, threads):
    for seed, axis, frac in [(1, 0, 0.5), (2, 1, 0.07), (3, 2, 0.93), (4, 0, 0.0), (5, 1, 1.5)]:  # incl. all-false / all-true
        assert lib.yh_host_partition_check(n, seed, threads, axis, frac, 0) == 0
    for seed, dup in [(6, 2), (7, 5)]:  # heavy duplicates of the key
        assert lib.yh_host_partition_check(n, seed, threads, 0, 0.5, dup) == 0
