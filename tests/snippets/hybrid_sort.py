"""Shared body of the hybrid-sort checks (partial LSD passes + segment fix-up, radix_sort.cu::segment_fix_kernel).
Run by tests/test_emu_kernels.py on the CPU emulator and by tests/test_sort_hybrid_gpu.py on the GPU, both with
B2_SORT_HYBRID_MIN=0 so that small inputs take the hybrid plan. `plc`, `np`, `osort`, `L` are provided by the caller;
SIZES scales the cases."""
CODE = r"""
rng = np.random.default_rng(7)
def check(keys, vals=None, order=0):
    kc = plc.Column.from_numpy(keys)
    so = plc.sorting.sorted_order(plc.Table([kc]), [order], []).to_numpy()[0]
    assert np.array_equal(so, osort.sorted_order([(keys, None)], [order])), ("sorted_order", keys.dtype, len(keys), order)
    if keys.dtype.kind in "iu":
        s = plc.sorting.sort(plc.Table([kc]), [order], []).columns()[0].to_numpy()[0]
        e = np.sort(keys, kind="stable"); e = e[::-1] if order else e
        assert np.array_equal(s, e), ("sort", keys.dtype, len(keys), order)
    if vals is not None:
        got = plc.sorting.sort_by_key(plc.Table([plc.Column.from_numpy(vals)]), plc.Table([kc]), [order], []).columns()[0].to_numpy()[0]
        ex = osort.sort_by_key([(vals, None)], [(keys, None)], [order])[0][0]
        assert np.array_equal(got, ex), ("sort_by_key", keys.dtype, vals.dtype, len(keys), order)
def launches(k):
    b = L.kernel_launch_count()
    plc.sorting.sorted_order(plc.Table([plc.Column.from_numpy(k)]), [0], [])
    return L.kernel_launch_count() - b
for n in SIZES:
    for order in (0, 1):
        # uniform 64-bit keys: few passes + fix-up of tiny segments
        check(rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64), rng.integers(0, 1 << 40, n).astype(np.int64), order)
        # 50 prefixes x 2^20 low values: segments of several rows
        check((rng.integers(0, 50, n).astype(np.int64) << 40) | rng.integers(0, 1 << 20, n).astype(np.int64), rng.standard_normal(n), order)
        # ~n/2 prefixes above bit 40: segments of about two rows, many of them straddling a warp's 32 rows
        check((rng.integers(0, max(1, n // 2), n).astype(np.int64) << 40) | rng.integers(0, 1 << 40, n).astype(np.int64), rng.integers(0, 1 << 40, n).astype(np.int64), order)
        # long runs of one value (duplicates longer than the fix-up window)
        check(rng.integers(0, 3, n).astype(np.int64) * (1 << 50) + 7, rng.integers(0, 100, n).astype(np.int32), order)
        # mostly one hot key + uniform rest
        hot = rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64); hot[rng.random(n) < 0.4] = 12345
        check(hot, rng.integers(0, 100, n).astype(np.int64), order)
        # top four bytes identical to each other: the independence estimate is wrong -> overflow -> full LSD rerun
        r = rng.integers(0, 256, n).astype(np.uint64)
        check((r * np.uint64(0x0101010100000000)) | rng.integers(0, 1 << 32, n).astype(np.uint64), rng.integers(0, 100, n).astype(np.int64), order)
        kf = rng.standard_normal(n) * 1e10
        kf[rng.integers(0, n, max(1, n // 10))] = np.nan
        kf[rng.integers(0, n, max(1, n // 10))] = -0.0
        check(kf, None, order)
n = SIZES[-1]
r = rng.integers(0, 256, n).astype(np.uint64)
a = launches(rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64))
b = launches((r * np.uint64(0x0101010100000000)) | rng.integers(0, 1 << 32, n).astype(np.uint64))
assert b > a + 8, (a, b)   # the second input reran the full sort
print('HYBRID_OK')
"""
