"""pylibcudf.partitioning twin: cudf::partition / hash_partition contracts (cpp/include/cudf/partitioning.hpp:58-145,
examples at :72-89). GPU (and emulator) only: the functions are thin compositions over the validated b2_partition."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_partition_gtest_vectors(plc):
    # cpp/tests/partitioning/partition_test.cpp:124-191 (Identity / Reverse / SinglePartition), fixed-width column only
    for vdt in (np.int32, np.int64, np.float64, np.int8):
        for mdt in (np.int8, np.int32, np.uint16, np.int64):
            first = np.array([0, 1, 2, 3, 4, 5], vdt)
            t = plc.Table([plc.Column.from_numpy(first)])
            out, offs = plc.partitioning.partition(t, plc.Column.from_numpy(np.array([0, 1, 2, 3, 4, 5], mdt)), 6)
            assert offs == [0, 1, 2, 3, 4, 5, 6] and out.columns()[0].to_numpy()[0].tolist() == first.tolist()
            first = np.array([0, 1, 3, 7, 5, 13], vdt)
            t = plc.Table([plc.Column.from_numpy(first)])
            out, offs = plc.partitioning.partition(t, plc.Column.from_numpy(np.array([5, 4, 3, 2, 1, 0], mdt)), 6)
            assert offs == [0, 1, 2, 3, 4, 5, 6] and out.columns()[0].to_numpy()[0].tolist() == [13, 5, 7, 3, 1, 0]
            out, offs = plc.partitioning.partition(t, plc.Column.from_numpy(np.zeros(6, mdt)), 1)
            assert offs == [0, 6] and sorted(out.columns()[0].to_numpy()[0].tolist()) == sorted(first.tolist())
    # partition_test.cpp:147-169 offsets for map {9, 2} and 12 partitions
    t = plc.Table([plc.Column.from_numpy(np.array([1, 2], np.int32), np.array([False, True]))])
    out, offs = plc.partitioning.partition(t, plc.Column.from_numpy(np.array([9, 2], np.int32)), 12)
    assert offs == [0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2]
    v, m = out.columns()[0].to_numpy()
    assert m.tolist() == [True, False] and v[0] == 2
    # partitioning.hpp:72-89: a partition nobody maps to is empty; rows keep their order inside a partition
    t = plc.Table([plc.Column.from_numpy(np.array([10, 20, 30, 40, 50], np.int32)), plc.Column.from_numpy(np.arange(5, dtype=np.float64))])
    out, offs = plc.partitioning.partition(t, plc.Column.from_numpy(np.array([3, 3, 0, 3, 0], np.int32)), 4)
    assert offs == [0, 2, 2, 2, 5]
    assert out.columns()[1].to_numpy()[0].tolist() == [2.0, 4.0, 0.0, 1.0, 3.0]
    # errors (partition_test.cpp:55-77): size mismatch, nulls in the map
    with pytest.raises(RuntimeError):
        plc.partitioning.partition(t, plc.Column.from_numpy(np.array([0, 1], np.int32)), 2)
    with pytest.raises(RuntimeError):
        plc.partitioning.partition(t, plc.Column.from_numpy(np.zeros(5, np.int32), np.array([1, 1, 0, 1, 1], bool)), 2)
    out, offs = plc.partitioning.partition(t, plc.Column.from_numpy(np.zeros(5, np.int32)), 0)
    assert offs == [0] and out.num_rows() == 0


@pytest.mark.parametrize("mdt", [np.int32, np.int8, np.uint16, np.int64])
def test_partition_random(plc, mdt):
    rng = np.random.default_rng(5)
    for n, P in [(1, 1), (5000, 7), (40_000, 200), (30_000, 5000)]:
        P = min(P, 100) if np.dtype(mdt).itemsize == 1 else P
        m = rng.integers(0, P, n).astype(mdt)
        vals = rng.integers(0, 1 << 40, n)
        t = plc.Table([plc.Column.from_numpy(vals)])
        out, offs = plc.partitioning.partition(t, plc.Column.from_numpy(m), P)
        order = np.argsort(m, kind="stable")
        assert np.array_equal(out.columns()[0].to_numpy()[0], vals[order])
        assert offs == np.concatenate([[0], np.cumsum(np.bincount(m.astype(np.int64), minlength=P))]).tolist()


def _check_hash_partition(plc, table_cols, key_idx, P, seed=0, identity=False):
    from oracle import partition as opart

    t = plc.Table([plc.Column.from_numpy(v, m) for v, m in table_cols])
    kw = {} if not identity else {"hash_function": plc.partitioning.HashId.HASH_IDENTITY}
    out, offs = plc.partitioning.hash_partition(t, key_idx, P, seed=seed, **kw)
    exp_cols, exp_offs = opart.hash_partition(table_cols, [table_cols[i] for i in key_idx], P, seed, identity)
    assert offs == exp_offs
    for c, (ev, em) in zip(out.columns(), exp_cols):
        gv, gm = c.to_numpy()
        em_ = np.ones(len(ev), bool) if em is None else em
        gm_ = np.ones(len(gv), bool) if gm is None else gm
        assert np.array_equal(gm_, em_)
        assert np.array_equal(gv[em_], ev[em_], equal_nan=ev.dtype.kind == "f")


def test_hash_partition_matches_libcudf_row_hash(plc):
    """Partition assignment = murmur3 row hash % P exactly as libcudf computes it (oracle/partition.py), rows stable inside a
    partition; every fixed-width key type, nulls, several key columns, seeds, identity hash, many partitions."""
    rng = np.random.default_rng(6)
    n = 20_000
    rows = np.arange(n, dtype=np.int32)
    for dt in (np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64, np.bool_):
        if dt == np.bool_:
            k = rng.random(n) < 0.5
        elif np.dtype(dt).kind == "f":
            k = (rng.standard_normal(n) * 3).round(1).astype(dt)
            k[::50] = np.nan; k[::70] = -0.0; k[1::70] = 0.0
        else:
            k = rng.integers(0, 100, n).astype(dt)
        for P in (1, 2, 13, 64):
            _check_hash_partition(plc, [(k, None), (rows, None)], [0], P)
        _check_hash_partition(plc, [(k, rng.random(n) < 0.8), (rows, None)], [0], 7, seed=619)
    a, b, c = rng.integers(0, 50, n).astype(np.int64), rng.integers(0, 9, n).astype(np.int16), rng.standard_normal(n).round(1)
    _check_hash_partition(plc, [(a, None), (b, rng.random(n) < 0.9), (c, None), (rows, None)], [0, 1, 2], 11)
    _check_hash_partition(plc, [(a, None), (b, None), (rows, None)], [1, 0], 16, seed=42)
    _check_hash_partition(plc, [(a, None), (rows, None)], [0], 5, identity=True)
    _check_hash_partition(plc, [(a, None), (b, None), (rows, None)], [0, 1], 3000)   # more than 256 partitions
    # hash_partition_test.cpp:76-141: empty inputs / no key columns / zero partitions -> empty table, P + 1 zero offsets
    t = plc.Table([plc.Column.from_numpy(a), plc.Column.from_numpy(rows)])
    out, offs = plc.partitioning.hash_partition(t, [], 3)
    assert out.num_rows() == 0 and out.num_columns() == 2 and offs == [0, 0, 0, 0]
    out, offs = plc.partitioning.hash_partition(t, [0], 0)
    assert out.num_rows() == 0 and offs == [0]
    e = plc.Table([plc.Column.from_numpy(np.empty(0, np.int64))])
    out, offs = plc.partitioning.hash_partition(e, [0], 4)
    assert out.num_rows() == 0 and offs == [0, 0, 0, 0, 0]
    with pytest.raises(IndexError):   # hash_partition_test.cpp:52-60 std::out_of_range
        plc.partitioning.hash_partition(t, [0, 5], 3)
    with pytest.raises(ValueError):   # :62-71 key table with another row count
        plc.partitioning.hash_partition(t, plc.Table([plc.Column.from_numpy(np.arange(3, dtype=np.int64))]), 3)


def test_dlpack_roundtrip(plc):
    """Column.__dlpack__ / from_dlpack (pylibcudf interop to_dlpack / from_dlpack): zero-copy both ways, nulls rejected."""
    import torch

    x = torch.arange(1000, dtype=torch.int64, device="cuda") * 3
    c = plc.Column.from_dlpack(x)
    assert c.size() == 1000 and c.type().id() == plc.TypeId.INT64
    out = plc.sorting.sort(plc.Table([c]), [plc.Order.DESCENDING], []).columns()[0]
    y = torch.from_dlpack(out)
    assert y.data_ptr() == out.data().ptr and bool((y == torch.flip(x, [0])).all())
    masked = plc.Column.from_numpy(np.arange(4, dtype=np.int32), np.array([1, 0, 1, 1], bool))
    with pytest.raises(ValueError):
        torch.from_dlpack(masked)
