"""Shared body of the partitioned-groupby checks (groupby.cu::pgb_agg_kernel: partition by the top byte of mix64(key),
shared-memory aggregation per partition chunk, merge into the global table). Run with B2_GROUPBY_PARTITION_ROWS=1 on the
CPU emulator (tests/test_emu_kernels.py) and on the GPU (tests/test_groupby_partitioned_gpu.py); B2_GROUPBY_SMEM_SLOTS=64
makes the shared table overflow so that rows spill to the global table. `cu`, `o`, `np`, `sort_groups`,
`assert_columns_equal` come from the caller; CASES = [(rows, groups), ...]."""
CODE = r"""
rng = np.random.default_rng(3)
def check(keys, reqs):
    gk, gr = sort_groups(*cu.groupby(keys, reqs)); ek, er = sort_groups(*o.groupby(keys, reqs))
    assert_columns_equal(gk[0], ek[0], what="keys")
    for q, (_, kinds) in enumerate(reqs):
        for j, kind in enumerate(kinds):
            assert_columns_equal(gr[q][j], er[q][j], rtol=1e-6 if gr[q][j][0].dtype == np.float32 else 1e-9, what=kind)
for n, G in CASES:
    k = rng.integers(0, G, n).astype(np.int64) * 1_000_003 - 5
    c = rng.integers(0, 5, n).astype(np.int32)
    for vdt in (np.float64, np.int64, np.int32, np.float32, np.uint32):
        v = (rng.standard_normal(n) * 100).astype(vdt) if np.dtype(vdt).kind != 'u' else rng.integers(0, 1000, n).astype(vdt)
        check([(k, None)], [((v, None), ["sum", "count"]), ((c, None), ["count"])])
        check([(k, None)], [((v, None), ["min", "max", "mean"])])
    check([(k, None)], [((c, None), ["count_all"])])
    check([(k.astype(np.uint64), None)], [((v, None), ["sum"])])
    # the reserved key value of the shared table (mix64(key) == ~0) and a hot key
    kk = k.copy(); kk[::7] = np.uint64(0x89a5850e63c5f8aa).astype(np.int64); kk[::3] = k[0]   # mix64(0x89a5850e63c5f8aa) == 2^64 - 1
    check([(kk, None)], [((c, None), ["sum", "count"])])
print('PGB_OK')
"""
