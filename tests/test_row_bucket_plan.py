"""Host logic of the partitioned row build (rowsort.cuh: RowBuckets / SubBuckets; g2n.cu: row_passes, plan_buckets), through the
C ABI's g2n_plan_row_buckets -- no device needed.  Stage 4 replaces scipy/_coo.py tocsr's counting sort by row
(builders.py:283, utils.py:55); how it is organised must never change the result, so the plan only has to be VALID:
every row in exactly one bucket / sub-bucket, fan-outs and sizes within what the kernels' shared memory holds."""
import ctypes as C
import random

import pytest


def _plan(lib, entries, rows, ent):
    out = (C.c_uint32 * 8)()
    rc = lib.g2n_plan_row_buckets(C.c_uint64(entries), C.c_uint64(rows), C.c_int(ent), out)
    assert rc == 0
    return list(out)


@pytest.fixture(scope="module")
def lib():
    from gfa2network_b200 import _capi

    lib = _capi.load()
    lib.g2n_plan_row_buckets.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint32)]
    lib.g2n_plan_row_buckets.restype = C.c_int
    return lib


def _check(p, entries, rows, ent):
    passes, bucketed, nb, sh1, two, nsub, sh2, cap = p
    assert 1 <= passes <= 64
    if not bucketed:
        assert nb == 0 and two == 0
        return
    assert 1 <= nb <= 128
    last = max(rows, 1) - 1
    assert (last >> sh1) + 1 == nb  # every row has a bucket, the last bucket is the last row's
    if two:
        fan_shift = sh1 - sh2
        assert 0 <= fan_shift <= 8 and sh2 <= 12  # <= 256 sub-buckets per bucket, <= 4096 rows per sub-bucket
        assert nsub == nb << fan_shift and nsub <= 128 * 256
        assert cap == (12288 if ent == 8 else 16384)


def test_named_configurations(lib):
    # C2: row arrays fit L2 -> one flat pass
    assert _plan(lib, 6_000_000, 1_000_000, 4)[:2] == [1, 0]
    # C4d / C3 / C5 slab of the 8-GPU build: partitioned on two levels, sub-buckets of about half the shared-memory capacity
    for entries, rows, ent in ((120_000_000, 20_000_000, 4), (120_000_000, 20_000_000, 8), (100_000_000, 12_500_000, 4)):
        p = _plan(lib, entries, rows, ent)
        _check(p, entries, rows, ent)
        assert p[1] == 1 and p[4] == 1
        avg_per_sub = entries / rows * (1 << p[6])
        assert avg_per_sub <= p[7] * 0.75, p


def test_random_sizes(lib):
    rng = random.Random(5)
    for _ in range(3000):
        rows = rng.choice([1, 2, 3, 17, 1000, 10 ** 5, 10 ** 6, 10 ** 7, 10 ** 8, 2 ** 31 - 1]) + rng.randrange(0, 1000)
        rows = min(rows, 2 ** 31 - 1)
        entries = min(int(rows * rng.choice([0.01, 1, 2, 6, 50, 3000])) + rng.randrange(0, 100), 2 ** 32 - 32)
        ent = rng.choice([4, 8])
        _check(_plan(lib, entries, rows, ent), entries, rows, ent)


def test_forced_passes(lib, monkeypatch):
    """G2N_DBG_ROWPASS (how the GPU tests force the partitioned build on small inputs) and the switches that select one level /
    the older row-range passes."""
    monkeypatch.setenv("G2N_DBG_ROWPASS", "5")
    p = _plan(lib, 180_000, 30_000, 4)
    _check(p, 180_000, 30_000, 4)
    assert p[1] == 1 and p[4] == 1
    monkeypatch.setenv("G2N_DBG_NOSUB", "1")
    assert _plan(lib, 180_000, 30_000, 4)[4] == 0
    monkeypatch.setenv("G2N_DBG_NOBUCKET", "1")
    p = _plan(lib, 180_000, 30_000, 4)
    assert p[0] == 5 and p[1] == 0
    assert lib.g2n_plan_row_buckets(C.c_uint64(1), C.c_uint64(1), C.c_int(3), (C.c_uint32 * 8)()) != 0
