"""world_size-2 and -4 gloo tests (CPU) of the host logic of the sharded sort / join: splitters, all-to-all-v sizes,
rank-ordered result.  The device primitives are replaced by a numpy twin (tests may use the oracle)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class NumpyOps:
    def sort_keys(self, t):
        return torch.from_numpy(np.sort(t.numpy(), kind="stable"))

    def partition(self, columns, key, mode, splitters, nparts):
        k = key.numpy()
        if mode == 0:
            b = np.searchsorted(splitters.numpy(), k, side="right") if splitters is not None and splitters.numel() else np.zeros(len(k), int)
        else:
            from oracle import datagen  # noqa: F401  (hash only needs to agree between the two sides)

            x = k.view(np.uint64).copy()
            with np.errstate(over="ignore"):
                x ^= x >> np.uint64(33); x *= np.uint64(0xff51afd7ed558ccd); x ^= x >> np.uint64(33)
                x *= np.uint64(0xc4ceb9fe1a85ec53); x ^= x >> np.uint64(33)
            b = (x % np.uint64(nparts)).astype(np.int64)
        order = np.argsort(b, kind="stable")
        counts = np.bincount(b, minlength=nparts)
        offs = np.concatenate([[0], np.cumsum(counts)]).astype(int).tolist()
        return [torch.from_numpy(c.numpy()[order]) for c in columns], offs

    def sort_by_key(self, values, keys):
        return torch.from_numpy(values.numpy()[np.argsort(keys.numpy(), kind="stable")])

    def reduce(self, col, kind):
        return getattr(col, kind)()

    def groupby_sum_count(self, keys, values):
        k, inv = np.unique(keys.numpy(), return_inverse=True)
        return torch.from_numpy(k), torch.from_numpy(np.bincount(inv, weights=values.numpy(), minlength=len(k))), torch.from_numpy(np.bincount(inv, minlength=len(k)))

    def groupby_merge(self, keys, sums, counts):
        k, inv = np.unique(keys.numpy(), return_inverse=True)
        return (torch.from_numpy(k), torch.from_numpy(np.bincount(inv, weights=sums.numpy(), minlength=len(k))),
                torch.from_numpy(np.bincount(inv, weights=counts.numpy(), minlength=len(k)).astype(np.int64)))

    def inner_join(self, left, right):
        from oracle import join as ojoin

        l, r = ojoin.inner_join([(left.numpy(), None)], [(right.numpy(), None)])
        return torch.from_numpy(l), torch.from_numpy(r)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cudf_b200_sharded_shim import sharded  # see test body: loaded without the CUDA library

        rng = np.random.default_rng(100 + rank)
        n = 20_000 + 1000 * rank
        keys = torch.from_numpy(rng.integers(-10**9, 10**9, n))
        out = sharded.sort_by_key_sharded(keys, keys, ops=NumpyOps(), samples_per_rank=256)
        vals = torch.from_numpy(rng.integers(0, 100, n))
        out2 = sharded.sort_by_key_sharded(vals, keys, ops=NumpyOps(), samples_per_rank=256)
        lk = torch.from_numpy(rng.integers(0, 5000, 3000))
        rk = torch.from_numpy(rng.integers(0, 5000, 2000))
        jl, jr = sharded.inner_join_sharded(lk, rk, ops=NumpyOps())
        gk = torch.from_numpy(rng.integers(0, 50, 4000))
        gv = torch.from_numpy(rng.standard_normal(4000))
        mk, ms, mc = sharded.groupby_sum_count_sharded(gk, gv, ops=NumpyOps())
        tot = sharded.reduce_sharded(gv, "sum", ops=NumpyOps())
        q.put((rank, keys.numpy(), out.numpy(), vals.numpy(), out2.numpy(), lk.numpy(), rk.numpy(), jl.numpy(), jr.numpy(),
               gk.numpy(), gv.numpy(), mk.numpy(), ms.numpy(), mc.numpy(), float(tot)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_sort_and_join_gloo(tmp_path, monkeypatch, world):
    # load cudf_b200/sharded.py as a standalone module: the package __init__ needs the CUDA library
    import importlib.util
    import sys
    import types
    from pathlib import Path

    shim_dir = tmp_path / "cudf_b200_sharded_shim"
    shim_dir.mkdir()
    src = (Path(__file__).resolve().parent.parent / "cudf_b200" / "sharded.py").read_text()
    (shim_dir / "__init__.py").write_text("")
    (shim_dir / "sharded.py").write_text(src)
    monkeypatch.setenv("PYTHONPATH", f"{tmp_path}{os.pathsep}{Path(__file__).resolve().parent.parent}{os.pathsep}" + os.environ.get("PYTHONPATH", ""))
    sys.path.insert(0, str(tmp_path))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_keys = np.concatenate([r[1] for r in res])
    got = np.concatenate([r[2] for r in res])  # rank-ordered concatenation is the global order
    assert np.array_equal(got, np.sort(all_keys))
    assert min(len(r[2]) for r in res) > 0.3 * len(all_keys) / world  # splitters balance the shards
    all_vals = np.concatenate([r[3] for r in res])
    got2 = np.concatenate([r[4] for r in res])
    order = np.argsort(all_keys, kind="stable")
    # ties between equal keys may interleave across source ranks: compare as multisets per key
    assert np.array_equal(np.sort(all_keys[order] * 1000 + all_vals[order]), np.sort(np.sort(all_keys) * 1000 + got2))
    # join: global row ids of both sides
    L = np.concatenate([r[5] for r in res]); R = np.concatenate([r[6] for r in res])
    jl = np.concatenate([r[7] for r in res]); jr = np.concatenate([r[8] for r in res])
    from oracle import join as ojoin

    el, er = ojoin.inner_join([(L, None)], [(R, None)])
    gl, gr = ojoin.canonical(jl, jr)
    assert np.array_equal(gl, el) and np.array_equal(gr, er)
    # groupby / reduce: every rank holds the merged result
    GK = np.concatenate([r[9] for r in res]); GV = np.concatenate([r[10] for r in res])
    uk, inv = np.unique(GK, return_inverse=True)
    for r in res:
        o = np.argsort(r[11])
        assert np.array_equal(r[11][o], uk)
        np.testing.assert_allclose(r[12][o], np.bincount(inv, weights=GV), rtol=1e-9)
        assert np.array_equal(r[13][o], np.bincount(inv))
        np.testing.assert_allclose(r[14], GV.sum(), rtol=1e-9)
