"""CPU model of the index arithmetic of cudf_b200/csrc/partition.cu.

The stable P-way partition computes, per 4096-row tile, bucket counts (bucket_kernel), a bucket-major
exclusive scan over (bucket, tile) (tile_scan_kernel), and a stable destination per row (dest_kernel).
The experimental staged scatter (scatter_to_staged_kernel) then regroups each tile by bucket in shared
memory and writes (tile, bucket) runs. This file restates that arithmetic in numpy and checks it against
a stable argsort, so that an indexing slip in the kernels' design shows up without a GPU.
"""
import numpy as np
import pytest

PT_TILE = 4096


def plan(ids, P):
    n = len(ids)
    ntiles = (n + PT_TILE - 1) // PT_TILE
    counts = np.zeros((ntiles, P), np.int64)
    for t in range(ntiles):
        counts[t] = np.bincount(ids[t * PT_TILE:(t + 1) * PT_TILE], minlength=P)
    # bucket-major exclusive scan: all tiles of bucket 0, then bucket 1, ...
    flat = counts.T.reshape(-1)
    starts = (np.cumsum(flat) - flat).reshape(P, ntiles).T.copy()
    totals = counts.sum(axis=0)
    bucket_start = np.concatenate([[0], np.cumsum(totals)])
    dest = np.empty(n, np.int64)
    for t in range(ntiles):
        run = starts[t].copy()
        for j in range(t * PT_TILE, min(n, (t + 1) * PT_TILE)):
            b = ids[j]
            dest[j] = run[b]
            run[b] += 1
    return starts, bucket_start, dest


def staged_scatter(vals, ids, P, starts, bucket_start, dest):
    n = len(ids)
    ntiles = starts.shape[0]
    outs = [np.full(bucket_start[b + 1] - bucket_start[b], -1, np.int64) for b in range(P)]
    for t in range(ntiles):
        row0 = t * PT_TILE
        rows = min(PT_TILE, n - row0)
        gbase = starts[t]
        g1 = starts[t + 1] if t + 1 < ntiles else bucket_start[1:]
        cnt = g1 - gbase
        soff = np.concatenate([[0], np.cumsum(cnt)])
        assert soff[-1] == rows
        stage = np.full(rows, -1, np.int64)
        sb = np.full(rows, 255, np.int64)
        for j in range(rows):
            r = row0 + j
            b = ids[r]
            p = soff[b] + (dest[r] - gbase[b])
            assert 0 <= p < rows and sb[p] == 255
            stage[p] = vals[r]
            sb[p] = b
        for q in range(rows):
            b = sb[q]
            outs[b][(gbase[b] - bucket_start[b]) + (q - soff[b])] = stage[q]
    return outs


@pytest.mark.parametrize("n,P", [(1, 1), (17, 2), (4096, 8), (4097, 8), (3 * 4096 + 5, 5), (20000, 128), (9000, 3)])
def test_partition_plan_and_staged_scatter(n, P):
    rng = np.random.default_rng(n * 131 + P)
    ids = rng.integers(0, P, n).astype(np.int64)
    if P > 2:
        ids[ids == 1] = 0  # an empty bucket
    vals = rng.integers(0, 1 << 40, n)
    starts, bucket_start, dest = plan(ids, P)
    order = np.argsort(ids, kind="stable")
    expect = np.empty(n, np.int64)
    expect[order] = np.arange(n)
    assert np.array_equal(dest, expect)
    outs = staged_scatter(vals, ids, P, starts, bucket_start, dest)
    got = np.concatenate(outs) if outs else np.empty(0, np.int64)
    assert np.array_equal(got, vals[order])
