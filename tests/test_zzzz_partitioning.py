"""pylibcudf.partitioning twin: cudf::partition / hash_partition contracts (cpp/include/cudf/partitioning.hpp:58-145,
examples at :72-89). GPU (and emulator) only: the functions are thin compositions over the validated b2_partition."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_partition_header_example(plc):
    # partitioning.hpp:72-80: t = {10, 20, 30, 40, 50}, map {0, 1, 0, 1, 0}, 2 partitions -> offsets {0, 3}
    t = plc.Table([plc.Column.from_numpy(np.array([10, 20, 30, 40, 50], np.int32)), plc.Column.from_numpy(np.arange(5, dtype=np.float64))])
    out, offs = plc.partitioning.partition(t, plc.Column.from_numpy(np.array([0, 1, 0, 1, 0], np.int32)), 2)
    assert offs == [0, 3]
    v = out.columns()[0].to_numpy()[0].tolist()
    assert sorted(v[:3]) == [10, 30, 50] and sorted(v[3:]) == [20, 40]
    assert out.columns()[1].to_numpy()[0].tolist() == [0.0, 2.0, 4.0, 1.0, 3.0]  # rows keep their order inside a partition
    # a partition nobody maps to is empty (partitioning.hpp:82-89)
    out, offs = plc.partitioning.partition(t, plc.Column.from_numpy(np.array([3, 3, 0, 3, 0], np.int32)), 4)
    assert offs == [0, 2, 2, 2]


@pytest.mark.parametrize("mdt", [np.int32, np.int8, np.uint16, np.int64])
def test_partition_random(plc, mdt):
    rng = np.random.default_rng(5)
    for n, P in [(1, 1), (5000, 7), (40_000, 200)]:
        P = min(P, 100) if np.dtype(mdt).itemsize == 1 else P
        m = rng.integers(0, P, n).astype(mdt)
        vals = rng.integers(0, 1 << 40, n)
        t = plc.Table([plc.Column.from_numpy(vals)])
        out, offs = plc.partitioning.partition(t, plc.Column.from_numpy(m), P)
        order = np.argsort(m, kind="stable")
        assert np.array_equal(out.columns()[0].to_numpy()[0], vals[order])
        assert offs == np.concatenate([[0], np.cumsum(np.bincount(m.astype(np.int64), minlength=P))[:-1]]).tolist()


def test_hash_partition_contract(plc):
    rng = np.random.default_rng(6)
    n, P = 30_000, 13
    k = rng.integers(0, 500, n)
    t = plc.Table([plc.Column.from_numpy(k), plc.Column.from_numpy(np.arange(n, dtype=np.int32))])
    out, offs = plc.partitioning.hash_partition(t, [0], P)
    ok, rows = out.columns()[0].to_numpy()[0], out.columns()[1].to_numpy()[0]
    assert np.array_equal(np.sort(rows), np.arange(n)) and np.array_equal(k[rows], ok)      # a permutation of the rows
    bounds = offs + [n]
    part_of = {}
    for p in range(P):
        for key in np.unique(ok[bounds[p]:bounds[p + 1]]):
            assert part_of.setdefault(int(key), p) == p                                       # equal keys -> same partition
    with pytest.raises(ValueError):
        plc.partitioning.hash_partition(t, [0, 1], P)


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
