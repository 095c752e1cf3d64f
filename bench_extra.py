"""Secondary measurements for bench.py --extra: hash inner_join (BASELINE configs[2]), groupby (configs[3]),
scan / reduce / segmented reduce.  Each entry reports rows/s, the algorithmic bytes of SURVEY §8d
and the achieved fraction of the measured HBM peak.  Data setup uses torch ops; the timed region is the library."""
from __future__ import annotations

import ctypes as C


def _fill(_lib, t, n, stream_id, kind=0, modulus=0, seed=0x5EED0001):
    _lib.check(_lib.lib.b2_fill_splitmix64(C.c_void_p(t.data_ptr()), n, seed, stream_id << 40, kind, modulus, _lib.stream_arg(None)))
    return t


def _time(torch, fn, steps=3, warmup=2):
    for _ in range(warmup):
        out = fn()
        del out
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
        del out
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run(plc, _lib, n, peak_gbs):
    import torch

    res = {}
    dev = "cuda"

    def entry(ms, rows, alg_bytes, **kw):
        gbs = alg_bytes / (ms / 1e3) / 1e9
        return dict(ms=ms, rows_per_s=rows / (ms / 1e3), algorithmic_bytes=alg_bytes, achieved_GBps=gbs, frac_of_peak=gbs / peak_gbs, **kw)

    # ---- scan / reduce / segmented reduce (int64 and float64) ----
    x = _fill(_lib, torch.empty(n, dtype=torch.int64, device=dev), n, 7)
    f = _fill(_lib, torch.empty(n, dtype=torch.float64, device=dev), n, 8, kind=1)
    ci, cf = plc.Column.from_torch(x), plc.Column.from_torch(f)
    agg = plc.aggregation
    res["scan_sum_int64"] = entry(_time(torch, lambda: plc.reduce.scan(ci, agg.sum(), plc.reduce.ScanType.INCLUSIVE)), n, 16 * n)
    res["scan_sum_float64"] = entry(_time(torch, lambda: plc.reduce.scan(cf, agg.sum(), plc.reduce.ScanType.INCLUSIVE)), n, 16 * n)
    res["reduce_sum_int64"] = entry(_time(torch, lambda: plc.reduce.reduce(ci, agg.sum(), plc.DataType(plc.TypeId.INT64))), n, 8 * n)
    res["reduce_sum_float64"] = entry(_time(torch, lambda: plc.reduce.reduce(cf, agg.sum(), plc.DataType(plc.TypeId.FLOAT64))), n, 8 * n)
    # keys-only radix: cudf::sort of one int64 column, and (opt-in B2_SORT_ALIAS=1) sort_by_key(T, T) routed to it
    ti = plc.Table([ci])
    res["sort_single_int64_keys_only"] = entry(_time(torch, lambda: plc.sorting.sort(ti, [plc.Order.ASCENDING], [])), n, 136 * n,
                                               note="histogram 8 B/row + 8 passes x 16 B/row")
    import os

    prev = os.environ.get("B2_SORT_ALIAS")
    os.environ["B2_SORT_ALIAS"] = "1"
    try:
        res["sort_by_key_aliased_opt_in"] = entry(_time(torch, lambda: plc.sorting.sort_by_key(ti, ti, [plc.Order.ASCENDING], [])), n, 136 * n)
    finally:
        if prev is None:
            del os.environ["B2_SORT_ALIAS"]
        else:
            os.environ["B2_SORT_ALIAS"] = prev
    del ti
    S = 1_000_000
    offs = torch.linspace(0, n, S + 1, device=dev).to(torch.int32)
    co = plc.Column.from_torch(offs)
    res["segmented_reduce_sum_float64_1e6_segments"] = entry(
        _time(torch, lambda: plc.reduce.segmented_reduce(cf, co, agg.sum(), plc.DataType(plc.TypeId.FLOAT64))), n, 8 * n + 4 * (S + 1) + 8 * S)
    del x, ci, offs, co

    # ---- groupby: int64 key with 1e6 groups, sum(float64) + count(int32) ----
    G = 1_000_000
    k = _fill(_lib, torch.empty(n, dtype=torch.int64, device=dev), n, 9, kind=2, modulus=G)
    v2 = _fill(_lib, torch.empty(n, dtype=torch.int32, device=dev), n, 10, kind=3)
    gb = plc.groupby.GroupBy(plc.Table([plc.Column.from_torch(k)]))
    reqs = [plc.groupby.GroupByRequest(cf, [agg.sum()]), plc.groupby.GroupByRequest(plc.Column.from_torch(v2), [agg.count()])]
    ms = _time(torch, lambda: gb.aggregate(reqs))
    keys_out, _ = gb.aggregate(reqs)
    res["groupby_sum_count_1e6_groups"] = entry(ms, n, 16 * n + 20 * G, groups=keys_out.num_rows(), atomics_per_s=2 * n / (ms / 1e3))
    del k, v2, gb, reqs, keys_out

    # ---- inner join: |R| = |L| = n, 10 % of probe rows match exactly once; payload gather with 50 % nulls ----
    try:
        rk = _fill(_lib, torch.empty(n, dtype=torch.int64, device=dev), n, 1)
        lk = _fill(_lib, torch.empty(n, dtype=torch.int64, device=dev), n, 6)          # fresh keys (match prob ~ 0)
        u = _fill(_lib, torch.empty(n, dtype=torch.float64, device=dev), n, 5, kind=1)
        sel = _fill(_lib, torch.empty(n, dtype=torch.int64, device=dev), n, 4, kind=2, modulus=n)
        hit = u < 0.10
        del u
        lk[hit] = rk[sel[hit]]
        del sel, hit
        L, R = plc.Table([plc.Column.from_torch(lk)]), plc.Table([plc.Column.from_torch(rk)])
        _lib.lib.b2_profile_reset()
        _lib.lib.b2_profile_enable(1)
        ms = _time(torch, lambda: plc.join.inner_join(L, R, plc.NullEquality.EQUAL), steps=2, warmup=1)
        li, ri = plc.join.inner_join(L, R, plc.NullEquality.EQUAL)
        M = li.size()
        slots = 1 << (2 * n - 1).bit_length()
        alg = 16 * slots + 24 * n + 24 * n + 8 * M
        res["inner_join_10pct"] = entry(ms, n, alg, matches=M, table_slots=slots,
                                        phases_ms={k_: _lib.profile_get(k_)[0] / max(_lib.profile_get(k_)[1], 1) for k_ in ("join_build", "join_count", "join_retrieve")})
        _lib.lib.b2_profile_enable(0)
        # opt-in partitioned shared-memory join (radix_join.cu) on the same inputs
        import os

        prev = os.environ.get("B2_JOIN_RADIX_ROWS")
        os.environ["B2_JOIN_RADIX_ROWS"] = "1000000"
        try:
            _lib.check(_lib.lib.b2_trim_pool())
            _lib.lib.b2_profile_reset()
            _lib.lib.b2_profile_enable(1)
            rms = _time(torch, lambda: plc.join.inner_join(L, R, plc.NullEquality.EQUAL), steps=2, warmup=1)
            li2, _ri2 = plc.join.inner_join(L, R, plc.NullEquality.EQUAL)
            res["inner_join_10pct_radix_opt_in"] = entry(rms, n, alg, matches=li2.size(), same_match_count=bool(li2.size() == M),
                                                         phases_ms={k_: _lib.profile_get(k_)[0] / max(_lib.profile_get(k_)[1], 1) for k_ in ("rjoin_partition", "rjoin_count", "rjoin_retrieve")})
            del li2, _ri2
        except Exception as ex:
            res["inner_join_10pct_radix_opt_in"] = {"error": repr(ex)[:200]}
        finally:
            _lib.lib.b2_profile_enable(0)
            if prev is None:
                del os.environ["B2_JOIN_RADIX_ROWS"]
            else:
                os.environ["B2_JOIN_RADIX_ROWS"] = prev
        # materialisation: gather both payload columns (float64, 50 % nulls) through the index columns
        pay = f
        nwords = (n + 31) // 32
        mask = _fill(_lib, torch.empty(nwords, dtype=torch.int32, device=dev), n, 3, kind=4)
        pcol = plc.Column.from_torch(pay, mask=mask)
        pt = plc.Table([pcol])
        gms = _time(torch, lambda: (plc.copying.gather(pt, li, plc.OutOfBoundsPolicy.DONT_CHECK), plc.copying.gather(pt, ri, plc.OutOfBoundsPolicy.DONT_CHECK)), steps=2, warmup=1)
        res["inner_join_materialise_payloads"] = entry(gms, M, 2 * M * (4 + 8 + 8) + 2 * M // 4, null_fraction=pcol.null_count() / n)
    except MemoryError as ex:  # not enough HBM for the 34 GB table next to the inputs
        res["inner_join_10pct"] = {"error": str(ex)[:200]}
    return res
