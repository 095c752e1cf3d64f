"""BASELINE.json full sizes (1e9 rows) on one B200: size-independent properties instead of an oracle comparison
(SURVEY §8c "Large-N parity").  Runs last (file name) because each case moves tens of GB."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu

N = 1_000_000_000


def _fill(_lib, t, n, stream_id, kind=0, modulus=0):
    _lib.check(_lib.lib.b2_fill_splitmix64(C.c_void_p(t.data_ptr()), n, 0x5EED0001, stream_id << 40, kind, modulus, _lib.stream_arg(None)))
    return t


def _enough_memory(torch, need_gb):
    from cudf_b200 import _lib

    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    _lib.check(_lib.lib.b2_trim_pool())
    free, _ = torch.cuda.mem_get_info()
    return free / 2**30 >= need_gb


def test_sort_by_key_1e9_properties(plc):
    import torch

    from cudf_b200 import _lib

    if not _enough_memory(torch, 70):
        pytest.skip("needs ~70 GB of free HBM")
    keys = _fill(_lib, torch.empty(N, dtype=torch.int64, device="cuda"), N, 0)
    t = plc.Table([plc.Column.from_torch(keys)])
    out = plc.sorting.sort_by_key(t, t, [plc.Order.ASCENDING], []).columns()[0].to_torch()
    assert out.numel() == N
    assert bool((out[1:] >= out[:-1]).all())                      # sortedness
    assert int(out.sum()) == int(keys.sum())                       # multiset preserved (wrap-around sums)
    assert int((out ^ (out >> 7)).sum()) == int((keys ^ (keys >> 7)).sum())
    del out
    order = plc.sorting.sorted_order(t, [plc.Order.DESCENDING], []).to_torch()
    assert int(order.long().sum()) == N * (N - 1) // 2             # a permutation of 0..N-1 (necessary condition)
    g = keys[order.long()[: 1 << 24]]
    assert bool((g[1:] <= g[:-1]).all())
    _lib.check(_lib.lib.b2_trim_pool())


def test_groupby_1e9_properties(plc):
    import torch

    from cudf_b200 import _lib

    if not _enough_memory(torch, 40):
        pytest.skip("needs ~40 GB of free HBM")
    G = 1_000_000
    k = _fill(_lib, torch.empty(N, dtype=torch.int64, device="cuda"), N, 9, kind=2, modulus=G)
    v = _fill(_lib, torch.empty(N, dtype=torch.float64, device="cuda"), N, 8, kind=1)
    gb = plc.groupby.GroupBy(plc.Table([plc.Column.from_torch(k)]))
    agg = plc.aggregation
    keys_out, res = gb.aggregate([plc.groupby.GroupByRequest(plc.Column.from_torch(v), [agg.sum(), agg.count()])])
    gk = keys_out.columns()[0].to_torch()
    sums, counts = res[0].columns()[0].to_torch(), res[0].columns()[1].to_torch()
    assert gk.numel() == G and int(gk.min()) == 0 and int(gk.max()) == G - 1
    assert int(torch.unique(gk).numel()) == G                      # every group exactly once
    assert int(counts.long().sum()) == N                           # counts are exact
    total = float(v.sum())
    assert abs(float(sums.sum()) - total) <= 1e-6 * abs(total)     # float sums: 1e-6 relative (north_star)
    _lib.check(_lib.lib.b2_trim_pool())


def test_inner_join_1e9_properties(plc):
    import torch

    from cudf_b200 import _lib

    if not _enough_memory(torch, 90):
        pytest.skip("needs ~90 GB of free HBM")
    rk = _fill(_lib, torch.empty(N, dtype=torch.int64, device="cuda"), N, 1)
    lk = _fill(_lib, torch.empty(N, dtype=torch.int64, device="cuda"), N, 6)
    u = _fill(_lib, torch.empty(N, dtype=torch.float64, device="cuda"), N, 5, kind=1)
    sel = _fill(_lib, torch.empty(N, dtype=torch.int64, device="cuda"), N, 4, kind=2, modulus=N)
    hit = u < 0.10
    del u
    nhit = int(hit.sum())
    lk[hit] = rk[sel[hit]]
    del sel, hit
    li, ri = plc.join.inner_join(plc.Table([plc.Column.from_torch(lk)]), plc.Table([plc.Column.from_torch(rk)]), plc.NullEquality.EQUAL)
    l, r = li.to_torch().long(), ri.to_torch().long()
    # every selected probe row matches (64-bit key collisions add a handful more), and every pair joins equal keys
    assert nhit <= l.numel() <= nhit + 1000
    assert bool((lk[l] == rk[r]).all())
    _lib.check(_lib.lib.b2_trim_pool())


# ---- oracle-exact comparisons at SURVEY §8c sizes (numpy on the host cores; tens of seconds each) ----------------------
def test_sort_1e8_exact_vs_numpy_stable(plc):
    """sorted_order and sort_by_key of 1e8 uniform int64 keys, bit-exact against np.argsort(kind='stable')."""
    import numpy as np
    import torch

    from cudf_b200 import _lib

    n = 100_000_000
    keys = _fill(_lib, torch.empty(n, dtype=torch.int64, device="cuda"), n, 0)
    vals = _fill(_lib, torch.empty(n, dtype=torch.float64, device="cuda"), n, 8, kind=1)
    kt = plc.Table([plc.Column.from_torch(keys)])
    order = plc.sorting.sorted_order(kt, [plc.Order.ASCENDING], []).to_torch().cpu().numpy()
    got = plc.sorting.sort_by_key(plc.Table([plc.Column.from_torch(vals)]), kt, [plc.Order.ASCENDING], []).columns()[0].to_torch().cpu().numpy()
    hk, hv = keys.cpu().numpy(), vals.cpu().numpy()
    exp = np.argsort(hk, kind="stable")
    assert order.dtype == np.int32 and np.array_equal(order, exp.astype(np.int32))
    assert np.array_equal(got, hv[exp])
    # a duplicate-heavy column (ties keep input order): keys mod 1000
    n = 30_000_000
    dk = _fill(_lib, torch.empty(n, dtype=torch.int64, device="cuda"), n, 9, kind=2, modulus=1000)
    order = plc.sorting.sorted_order(plc.Table([plc.Column.from_torch(dk)]), [plc.Order.DESCENDING], []).to_torch().cpu().numpy()
    hd = dk.cpu().numpy()
    exp = np.argsort(-hd, kind="stable")
    assert np.array_equal(order, exp.astype(np.int32))
    _lib.check(_lib.lib.b2_trim_pool())


def test_inner_join_2e7_canonical_pairs_exact(plc):
    """BASELINE configs[2] shape at 2e7 x 2e7 rows (above the 2^24-row limit of the partitioned path): canonical-sorted (left, right) pairs equal the oracle's, for the default
    path choice and for each join path forced (the oracle runs once)."""
    import os

    import numpy as np
    import torch

    from cudf_b200 import _lib
    from oracle import join as ojoin

    n = 20_000_000
    rk = _fill(_lib, torch.empty(n, dtype=torch.int64, device="cuda"), n, 1)
    lk = _fill(_lib, torch.empty(n, dtype=torch.int64, device="cuda"), n, 6)
    u = _fill(_lib, torch.empty(n, dtype=torch.float64, device="cuda"), n, 5, kind=1)
    sel = _fill(_lib, torch.empty(n, dtype=torch.int64, device="cuda"), n, 4, kind=2, modulus=n)
    hit = u < 0.10
    lk[hit] = rk[sel[hit]]
    lk[::1000] = lk[7]  # a probe-side hot key as well
    exp = ojoin.inner_join([(lk.cpu().numpy(), None)], [(rk.cpu().numpy(), None)])
    L, R = plc.Table([plc.Column.from_torch(lk)]), plc.Table([plc.Column.from_torch(rk)])
    prev = os.environ.get("B2_JOIN_RADIX_ROWS")
    try:
        for path in ("default", "hash_table", "partitioned"):
            if path == "default":
                os.environ.pop("B2_JOIN_RADIX_ROWS", None)
            else:
                os.environ["B2_JOIN_RADIX_ROWS"] = "0" if path == "hash_table" else "1000000"
            li, ri = plc.join.inner_join(L, R, plc.NullEquality.EQUAL)
            got = ojoin.canonical(li.to_torch().cpu().numpy(), ri.to_torch().cpu().numpy())
            assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]), path
    finally:
        if prev is None:
            os.environ.pop("B2_JOIN_RADIX_ROWS", None)
        else:
            os.environ["B2_JOIN_RADIX_ROWS"] = prev
    _lib.check(_lib.lib.b2_trim_pool())


def test_groupby_1e9_per_group_exact(plc):
    """BASELINE configs[3] at full size: every group's COUNT bit-exact and SUM(float64) within 1e-9 relative of np.bincount."""
    import numpy as np
    import torch

    from cudf_b200 import _lib

    if not _enough_memory(torch, 60):
        pytest.skip("needs ~60 GB of free HBM")
    G = 1_000_000
    k = _fill(_lib, torch.empty(N, dtype=torch.int64, device="cuda"), N, 9, kind=2, modulus=G)
    v = _fill(_lib, torch.empty(N, dtype=torch.float64, device="cuda"), N, 8, kind=1)
    c = _fill(_lib, torch.empty(N, dtype=torch.int32, device="cuda"), N, 10, kind=3)
    gb = plc.groupby.GroupBy(plc.Table([plc.Column.from_torch(k)]))
    agg = plc.aggregation
    keys_out, res = gb.aggregate([plc.groupby.GroupByRequest(plc.Column.from_torch(v), [agg.sum()]),
                                  plc.groupby.GroupByRequest(plc.Column.from_torch(c), [agg.count()])])
    gk = keys_out.columns()[0].to_torch().cpu().numpy()
    gs = res[0].columns()[0].to_torch().cpu().numpy()
    gc = res[1].columns()[0].to_torch().cpu().numpy()
    hk, hv = k.cpu().numpy(), v.cpu().numpy()
    del k, v, c
    order = np.argsort(gk, kind="stable")
    assert np.array_equal(gk[order], np.arange(G, dtype=np.int64))
    assert gc.dtype == np.int32 and np.array_equal(gc[order], np.bincount(hk, minlength=G).astype(np.int32))
    np.testing.assert_allclose(gs[order], np.bincount(hk, weights=hv, minlength=G), rtol=1e-9)
    _lib.check(_lib.lib.b2_trim_pool())
