"""b2_range_partition_counts / _scatter: the sharded sort's fused partition + exchange, exercised on ONE device with local destination
buffers (the peer case only changes the pointers). Stable range partition: bucket b = number of splitters <= key."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(plc, keys, vals, splitters, P):
    from cudf_b200 import _lib

    lib = _lib.lib
    kc = plc.Column.from_numpy(keys)
    sc = plc.Column.from_numpy(splitters) if P > 1 and splitters is not None else None
    sp = C.c_void_p(sc._data) if sc is not None else None
    counts = (C.c_int64 * P)()
    kv = kc._view()
    _lib.check(lib.b2_range_partition_counts(C.byref(kv), sp, P, _lib.stream_arg(None), counts))
    counts = list(counts)
    n = len(keys)
    kout = plc.Column.from_numpy(np.zeros(n + 8, keys.dtype))
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    kd = (C.c_void_p * P)(*[kout._data + int(offs[b]) * 8 for b in range(P)])
    vc = vout = None
    vd = None
    vv = None
    if vals is not None:
        vc = plc.Column.from_numpy(vals)
        vout = plc.Column.from_numpy(np.zeros(n + 8, vals.dtype))
        vd = (C.c_void_p * P)(*[vout._data + int(offs[b]) * vals.dtype.itemsize for b in range(P)])
        vv = vc._view()
    _lib.check(lib.b2_range_partition_scatter(C.byref(kv), C.byref(vv) if vv is not None else None, sp, P, kd, vd, _lib.stream_arg(None)))
    gk = kout.to_numpy()[0][:n]
    gv = vout.to_numpy()[0][:n] if vout is not None else None
    return counts, gk, gv


@pytest.mark.parametrize("kdt", [np.int64, np.uint64])
def test_range_partition_matches_stable_numpy(plc, kdt):
    rng = np.random.default_rng(31)
    for n in (1, 100, 6143, 6145, 50_001):
        for P in (1, 2, 8, 200):
            info = np.iinfo(kdt)
            keys = rng.integers(info.min, info.max, n, dtype=kdt, endpoint=True)
            if n > 10:
                keys[::7] = keys[3]                      # duplicates straddling a splitter value
            splitters = np.sort(rng.choice(keys, size=P - 1, replace=True)) if P > 1 else np.empty(0, kdt)
            bucket = np.searchsorted(splitters, keys, side="right") if P > 1 else np.zeros(n, int)
            order = np.argsort(bucket, kind="stable")
            exp_counts = np.bincount(bucket, minlength=P).tolist()
            for vdt in (None, np.int64, np.float32):
                vals = None if vdt is None else rng.integers(0, 1 << 30, n).astype(vdt)
                counts, gk, gv = _run(plc, keys, vals, splitters, P)
                assert counts == exp_counts, (n, P)
                assert np.array_equal(gk, keys[order]), (n, P, vdt)
                if vals is not None:
                    assert np.array_equal(gv, vals[order]), (n, P, vdt)


@pytest.mark.parametrize("kdt", [np.int64, np.uint64])
def test_hash_mode_matches_oracle_bucket(plc, kdt):
    """splitters == NULL: the sharded join's shuffle — stable partition by oracle.partition.shuffle_bucket."""
    from oracle.partition import shuffle_bucket

    rng = np.random.default_rng(77)
    for n in (1, 6143, 30_001):
        for P in (2, 3, 8):
            info = np.iinfo(kdt)
            keys = rng.integers(info.min, info.max, n, dtype=kdt, endpoint=True)
            keys[::5] = keys[0]
            bucket = shuffle_bucket(keys, P)
            order = np.argsort(bucket, kind="stable")
            vals = np.arange(n, dtype=np.int64)
            counts, gk, gv = _run(plc, keys, vals, None, P)
            assert counts == np.bincount(bucket, minlength=P).tolist(), (n, P)
            assert np.array_equal(gk, keys[order]) and np.array_equal(gv, order), (n, P)
