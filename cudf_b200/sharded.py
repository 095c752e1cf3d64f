"""Sharded (multi-GPU) sort_by_key and inner_join: one process per GPU, torch.distributed for the plumbing.

No libcudf equivalent (libcudf is single-GPU); this replaces the role of the rapidsmpf/dask shuffle above it
(python/cudf_polars/cudf_polars/streaming/sort.py:169-235 sample -> allgather -> splitters -> range partition
-> shuffle -> local sort; hash_partition -> shuffle -> local join for joins, cpp/include/cudf/partitioning.hpp).

  sort:  regular sample of each shard -> all_gather -> P-1 splitters -> b2_partition(range) ->
         all_to_all_single of the buckets (NCCL over NVLink) -> local LSD radix sort.
         Rank r ends up with the r-th key range; concatenating the shards in rank order is the global order.
  join:  b2_partition(hash) of (key, global row id) on both sides -> two all_to_all -> local hash join ->
         global row-id pairs (the output stays sharded).

The device work goes through an `ops` object (CudaOps = the CUDA library); tests inject a numpy twin to cover the
host logic with the gloo backend on CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import torch
import torch.distributed as dist

_TIMING = os.environ.get("B2_SHARD_TIMING", "0") == "1"
last_phases_ms: dict = {}   # phases of the most recent sharded call on this rank (filled when timing is on)


def enable_phase_timing(on: bool = True):
    """bench.py switches this on: phases are bracketed by CUDA events on the current stream (no extra synchronisation
    inside the step); `last_phases_ms` is resolved by phases() after the caller has synchronised."""
    global _TIMING
    _TIMING = bool(on)


class _Phase:
    """Phase timing of a sharded call: CUDA events on the current stream when a GPU is present, wall clock on CPU."""

    def __init__(self):
        self.marks = []

    def mark(self, name):
        if not _TIMING:
            return
        if torch.cuda.is_available():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))
        else:
            self.marks.append((name, time.perf_counter()))

    def done(self, tag):
        self.mark("_end")
        if not _TIMING:
            return
        global _pending
        _pending = (tag, self.marks)


_pending = None


def phases() -> dict:
    """Resolve the phase times of the last sharded call (synchronises the device)."""
    global _pending, last_phases_ms
    if _pending is None:
        return last_phases_ms
    tag, marks = _pending
    _pending = None
    out = {}
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    for (name, a), (_, b) in zip(marks[:-1], marks[1:]):
        ms = a.elapsed_time(b) if torch.cuda.is_available() else (b - a) * 1e3
        out[name] = out.get(name, 0.0) + ms
    last_phases_ms = out
    if os.environ.get("B2_SHARD_TIMING", "0") == "1" and dist.get_rank() == 0:
        print(f"[{tag}] " + " ".join(f"{k}={v:.1f}ms" for k, v in out.items()), flush=True)
    return out


class CudaOps:
    """Device primitives backed by libcudf_b200 (through the pylibcudf-named shim)."""

    def __init__(self):
        from . import _lib
        from . import pylibcudf as plc

        self.plc, self._lib = plc, _lib

    def sort_keys(self, t: torch.Tensor) -> torch.Tensor:
        plc = self.plc
        out = plc.sorting.sort(plc.Table([plc.Column.from_torch(t)]), [plc.Order.ASCENDING], [])
        return out.columns()[0].to_torch()

    def partition(self, columns, key, mode: int, splitters, nparts: int):
        """-> (list of partitioned tensors, offsets list[nparts+1]); stable within a bucket."""
        plc, lib = self.plc, self._lib
        tbl = plc.Table([plc.Column.from_torch(c) for c in columns])
        kcol = plc.Column.from_torch(key)
        tv, kv = tbl._view(), kcol._view()
        out = C.c_void_p()
        offs = (C.c_int32 * (nparts + 1))()
        sp = C.c_void_p(splitters.data_ptr()) if splitters is not None and splitters.numel() else None
        lib.check(lib.lib.b2_partition(C.byref(tv), C.byref(kv), mode, sp, nparts, lib.stream_arg(None), C.byref(out), offs))
        res = plc.Table._from_handle(out.value)
        return [c.to_torch() for c in res.columns()], list(offs)

    def partition_exchange(self, columns, key: torch.Tensor, mode: int, splitters, group=None, slot_base: int = 0, variant: str = "staged"):
        """Fused partition + all-to-all of several columns over peer memory: ONE plan (bucket + stable in-bucket rank of every
        row), then one scatter kernel per column that writes each row straight into its destination GPU's receive buffer
        (variant "staged": per-peer runs staged in shared memory first; "plain": row by row). Returns this rank's received
        columns (views of the exchange buffers slot_base.., valid until the next call that uses the same slots)."""
        plc, lib = self.plc, self._lib
        single = isinstance(columns, torch.Tensor)
        cols = [columns] if single else list(columns)
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        kcol = plc.Column.from_torch(key)
        kv = kcol._view()
        plan = C.c_void_p()
        counts = (C.c_int64 * world)()
        sp = C.c_void_p(splitters.data_ptr()) if splitters is not None and splitters.numel() else None
        lib.check(lib.lib.b2_partition_plan_create(C.byref(kv), mode, sp, world, lib.stream_arg(None), C.byref(plan), counts))
        try:
            mine = torch.tensor(list(counts), dtype=torch.int64, device=key.device)
            allc = torch.empty(world * world, dtype=torch.int64, device=key.device)
            dist.all_gather_into_tensor(allc, mine, group=group)
            cm = allc.view(world, world).cpu()                      # cm[r][d] = rows rank r sends to rank d
            recv_total = int(cm[:, rank].sum())
            max_recv = int(cm.sum(dim=0).max())
            max_send = int(cm.sum(dim=1).max())                    # every rank sees the same matrix: the capacity is agreed
            exs = [PeerExchange.get(lib, int(max(max_recv, max_send) * c.element_size() * 1.05) + (1 << 20), group, slot=slot_base + j)
                   for j, c in enumerate(cols)]
            my_off = cm[:rank, :].sum(dim=0)                          # rows written before mine in each destination
            dist.barrier(group=group)                                 # peers are done reading the previous contents
            scatter = lib.lib.b2_partition_scatter_staged if variant == "staged" else lib.lib.b2_partition_scatter
            for ex, c in zip(exs, cols):
                esz = c.element_size()
                dest = (C.c_void_p * world)(*[ex.peer_ptrs[d] + int(my_off[d]) * esz for d in range(world)])
                ccol = plc.Column.from_torch(c)
                cv = ccol._view()
                lib.check(scatter(plan, C.byref(cv), dest, lib.stream_arg(None)))
            dist.barrier(group=group)                                 # stream-ordered after the scatters: all buckets landed
            out = [ex.view(recv_total, c.dtype) for ex, c in zip(exs, cols)]
            return out[0] if single else out
        finally:
            lib.lib.b2_partition_plan_free(plan)

    def range_exchange(self, keys: torch.Tensor, values, splitters, group=None, slot_base: int = 0):
        """Range partition fused into ONE one-sweep pass that writes every bucket straight into its destination GPU's receive
        buffer (b2_range_partition_counts / _scatter; int64 / uint64 keys, optionally one 4- / 8-byte payload column).
        Returns (received keys, received values or None): views of the exchange buffers."""
        plc, lib = self.plc, self._lib
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        kcol = plc.Column.from_torch(keys)
        kv = kcol._view()
        sp = C.c_void_p(splitters.data_ptr()) if splitters is not None and splitters.numel() else None
        counts = (C.c_int64 * world)()
        lib.check(lib.lib.b2_range_partition_counts(C.byref(kv), sp, world, lib.stream_arg(None), counts))
        mine = torch.tensor(list(counts), dtype=torch.int64, device=keys.device)
        allc = torch.empty(world * world, dtype=torch.int64, device=keys.device)
        dist.all_gather_into_tensor(allc, mine, group=group)
        cm = allc.view(world, world).cpu()                          # cm[r][d] = rows rank r sends to rank d
        recv_total = int(cm[:, rank].sum())
        cap_rows = max(int(cm.sum(dim=0).max()), int(cm.sum(dim=1).max()))   # from the shared matrix: every rank agrees
        cols = [keys] + ([values] if values is not None else [])
        exs = [PeerExchange.get(lib, int(cap_rows * c.element_size() * 1.05) + (1 << 20), group, slot=slot_base + j) for j, c in enumerate(cols)]
        my_off = cm[:rank, :].sum(dim=0)                              # rows written before mine in each destination
        kd = (C.c_void_p * world)(*[exs[0].peer_ptrs[d] + int(my_off[d]) * keys.element_size() for d in range(world)])
        vd, vv = None, None
        if values is not None:
            vd = (C.c_void_p * world)(*[exs[1].peer_ptrs[d] + int(my_off[d]) * values.element_size() for d in range(world)])
            vcol = plc.Column.from_torch(values)
            vv = vcol._view()
        dist.barrier(group=group)                                     # peers are done reading the previous contents
        lib.check(lib.lib.b2_range_partition_scatter(C.byref(kv), C.byref(vv) if vv is not None else None, sp, world, kd, vd, lib.stream_arg(None)))
        dist.barrier(group=group)                                     # stream-ordered after the pass: all buckets landed
        rk = exs[0].view(recv_total, keys.dtype)
        return rk, (exs[1].view(recv_total, values.dtype) if values is not None else None)

    def sort_by_key(self, values: torch.Tensor, keys: torch.Tensor) -> torch.Tensor:
        plc = self.plc
        out = plc.sorting.sort_by_key(plc.Table([plc.Column.from_torch(values)]), plc.Table([plc.Column.from_torch(keys)]),
                                      [plc.Order.ASCENDING], [])
        return out.columns()[0].to_torch()

    def reduce(self, col: torch.Tensor, kind: str) -> torch.Tensor:
        plc = self.plc
        c = plc.Column.from_torch(col)
        agg = {"sum": plc.aggregation.sum, "min": plc.aggregation.min, "max": plc.aggregation.max}[kind]()
        s = plc.reduce.reduce(c, agg, c.type())
        return torch.tensor(s.to_py(), dtype=col.dtype, device=col.device)

    def groupby_sum_count(self, keys: torch.Tensor, values: torch.Tensor):
        plc = self.plc
        gb = plc.groupby.GroupBy(plc.Table([plc.Column.from_torch(keys)]))
        k, res = gb.aggregate([plc.groupby.GroupByRequest(plc.Column.from_torch(values), [plc.aggregation.sum(), plc.aggregation.count()])])
        return k.columns()[0].to_torch(), res[0].columns()[0].to_torch(), res[0].columns()[1].to_torch()

    def groupby_merge(self, keys: torch.Tensor, sums: torch.Tensor, counts: torch.Tensor):
        plc = self.plc
        gb = plc.groupby.GroupBy(plc.Table([plc.Column.from_torch(keys)]))
        k, res = gb.aggregate([plc.groupby.GroupByRequest(plc.Column.from_torch(sums), [plc.aggregation.sum()]),
                               plc.groupby.GroupByRequest(plc.Column.from_torch(counts), [plc.aggregation.sum()])])
        return k.columns()[0].to_torch(), res[0].columns()[0].to_torch(), res[1].columns()[0].to_torch()

    def inner_join(self, left: torch.Tensor, right: torch.Tensor):
        plc = self.plc
        l, r = plc.join.inner_join(plc.Table([plc.Column.from_torch(left)]), plc.Table([plc.Column.from_torch(right)]),
                                   plc.NullEquality.EQUAL)
        return l.to_torch(), r.to_torch()


class PeerExchange:
    """Receive buffers mapped into every peer with CUDA IPC (b2_ipc_*): the partition scatter kernel writes each
    bucket directly into its destination GPU's buffer over NVLink, so partition + all-to-all is ONE kernel
    (plus a barrier) instead of scatter -> NCCL send/recv -> copy.  One instance per (group, capacity); reused
    across calls."""

    _cache: dict = {}

    def __init__(self, lib, capacity_bytes: int, group=None):
        self._lib, self.capacity, self.group = lib, int(capacity_bytes), group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        ptr = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        lib.check(lib.lib.b2_ipc_alloc(self.capacity, C.byref(ptr), handle))
        self.local_ptr = ptr.value
        mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine, group=group)
        self.peer_ptrs = []
        for r in range(world):
            if r == rank:
                self.peer_ptrs.append(self.local_ptr)
                continue
            hb = (C.c_uint8 * 64)(*allh[r].cpu().tolist())
            p = C.c_void_p()
            lib.check(lib.lib.b2_ipc_open(hb, C.byref(p)))
            self.peer_ptrs.append(p.value)
        dist.barrier(group=group)

    @classmethod
    def get(cls, lib, capacity_bytes: int, group=None, slot: int = 0) -> "PeerExchange":
        """`capacity_bytes` must be computed from values every rank agrees on (the all-gathered count matrix): the
        constructor runs collectives, so all ranks have to take the same branch here."""
        key = (id(group), dist.get_world_size(group), slot)
        cur = cls._cache.get(key)
        if cur is None or cur.capacity < capacity_bytes:
            if cur is not None:
                cur.close()
            cur = cls(lib, capacity_bytes, group)
            cls._cache[key] = cur
        return cur

    def close(self):
        """Unmap the peers' buffers and free ours (all ranks call this together, before the replacement is built)."""
        rank = dist.get_rank(self.group)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # nobody still writes into the old buffers
        for r, p in enumerate(self.peer_ptrs):
            if r != rank and p:
                self._lib.lib.b2_ipc_close(C.c_void_p(p))
        dist.barrier(group=self.group)  # every mapping is closed before the owner frees
        self._lib.lib.b2_ipc_free(C.c_void_p(self.local_ptr))
        self.peer_ptrs, self.local_ptr = [], None

    def view(self, nelems: int, dtype: torch.dtype) -> torch.Tensor:
        import numpy as np

        from .pylibcudf.column import DeviceSpan

        npdt = np.dtype(str(dtype).replace("torch.", ""))
        return torch.as_tensor(DeviceSpan(self.local_ptr, nelems, npdt, self), device="cuda")


def _exchange(buckets: torch.Tensor, offsets, group=None) -> torch.Tensor:
    """all-to-all-v of contiguous buckets; returns the concatenation of what every rank sent to us."""
    world = dist.get_world_size(group)
    send = torch.tensor([offsets[i + 1] - offsets[i] for i in range(world)], dtype=torch.int64, device=buckets.device)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    send_l, recv_l = send.tolist(), recv.tolist()
    out = torch.empty(sum(recv_l), dtype=buckets.dtype, device=buckets.device)
    dist.all_to_all_single(out, buckets, output_split_sizes=recv_l, input_split_sizes=send_l, group=group)
    return out


def _exchange_peer(cols, offsets, lib, group=None, slot_base: int = 0):
    """all-to-all-v of the contiguous buckets of several columns over peer memory: every rank copies bucket d straight into
    rank d's receive buffer (CUDA-IPC mapped, b2_peer_copy: one contiguous run per destination, so NVLink sees large
    sequential writes). Returns the received columns (views of the exchange buffers of slots slot_base.., valid until
    the next exchange that uses the same slot)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = cols[0].device
    mine = torch.tensor([offsets[i + 1] - offsets[i] for i in range(world)], dtype=torch.int64, device=dev)
    allc = torch.empty(world * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allc, mine, group=group)
    cm = allc.view(world, world).cpu()                          # cm[r][d] = rows rank r sends to rank d
    recv_total = int(cm[:, rank].sum())
    cap_rows = max(int(cm.sum(dim=0).max()), int(cm.sum(dim=1).max()))
    my_off = cm[:rank, :].sum(dim=0)                              # rows written before mine in each destination
    exs = [PeerExchange.get(lib, int(cap_rows * c.element_size() * 1.05) + (1 << 20), group, slot=slot_base + j) for j, c in enumerate(cols)]
    dist.barrier(group=group)                                     # peers are done reading the previous contents
    st = lib.stream_arg(None)
    for j, c in enumerate(cols):
        esz = c.element_size()
        for step in range(world):                                 # staggered: at any time every rank writes to a different peer
            d = (rank + step) % world
            cnt = int(cm[rank][d])
            if cnt:
                lib.check(lib.lib.b2_peer_copy(C.c_void_p(exs[j].peer_ptrs[d] + int(my_off[d]) * esz),
                                               C.c_void_p(c.data_ptr() + int(offsets[d]) * esz), cnt * esz, st))
    dist.barrier(group=group)                                     # stream-ordered after the copies: all buckets landed
    return [ex.view(recv_total, c.dtype) for ex, c in zip(exs, cols)]


def _exchange_cols(cols, offsets, ops, group=None, slot_base: int = 0):
    """Bucket exchange of partitioned columns: peer copies on GPUs (B2_SHARD_XCHG=nccl selects all_to_all_single), the
    process group's all-to-all on CPU (gloo tests)."""
    if cols[0].is_cuda and hasattr(ops, "_lib") and os.environ.get("B2_SHARD_XCHG", "peer") != "nccl":
        return _exchange_peer(cols, offsets, ops._lib, group, slot_base)
    return [_exchange(c, offsets, group) for c in cols]


def _exchange_variant(t: torch.Tensor, ops) -> str:
    """Bucket exchange on GPUs (B2_SHARD_P2P forces one; measurements in profiles/r2_multi_gpu.md):
      "fused" (default on int64 keys): the partition is ONE one-sweep pass whose digit is the destination rank (sort: number of
          splitters <= key; join: a hash of the key) and whose per-(tile, bucket) runs are written straight into the peers'
          receive buffers (b2_range_partition_*; 2 GPUs: 9.2 ms per 1e9 keys)
      "staged": plan (bucket + stable rank per row) + scatter kernel, per-peer runs of a 4096-row tile staged in shared memory
          before the remote stores (2 GPUs: 14.0 ms per 1e9 rows, 8 GPUs: 22.1 ms)
      "1" / "plain": fused, row-by-row remote stores (2 GPUs: 16.9 ms; at 8 GPUs a warp's rows split into 32-byte writes)
      "0" / "copy": b2_partition, then one contiguous b2_peer_copy per destination (B2_SHARD_XCHG=nccl: all_to_all_single)
    CPU tensors (gloo tests) always take the partition + process-group all-to-all path."""
    if not (t.is_cuda and hasattr(ops, "partition_exchange")):
        return "copy"
    env = os.environ.get("B2_SHARD_P2P", "")
    if env in ("0", "copy"):
        return "copy"
    if env in ("1", "plain"):
        return "plain"
    if env == "staged":
        return "staged"
    return "fused"


def choose_splitters(samples_sorted: torch.Tensor, world: int) -> torch.Tensor:
    """P-1 splitters at the i/P quantiles of the gathered, sorted sample."""
    m = samples_sorted.numel()
    idx = torch.tensor([(i * m) // world for i in range(1, world)], dtype=torch.int64, device=samples_sorted.device)
    return samples_sorted[idx.clamp_(0, max(m - 1, 0))] if m else samples_sorted[:0]


def sort_by_key_sharded(values: torch.Tensor, keys: torch.Tensor, ops=None, group=None, samples_per_rank: int = 1 << 16) -> torch.Tensor:
    """Global sort of the row-sharded (values, keys); returns this rank's slice of the result (values ordered by key)."""
    ops = ops or CudaOps()
    world = dist.get_world_size(group)
    if world == 1:
        return ops.sort_by_key(values, keys)
    ph = _Phase()
    ph.mark("sample")
    n = keys.numel()
    stride = max(1, n // samples_per_rank)
    sample = keys[::stride][:samples_per_rank].contiguous()
    # every rank contributes exactly samples_per_rank entries (pad by repeating the last sample)
    if sample.numel() < samples_per_rank:
        pad = sample[-1:].expand(samples_per_rank - sample.numel()) if sample.numel() else keys.new_zeros(samples_per_rank)
        sample = torch.cat([sample, pad])
    gathered = torch.empty(world * samples_per_rank, dtype=keys.dtype, device=keys.device)
    dist.all_gather_into_tensor(gathered, sample, group=group) if keys.is_cuda else dist.all_gather(
        list(gathered.view(world, -1).unbind(0)), sample, group=group)
    splitters = choose_splitters(ops.sort_keys(gathered), world)
    same = values.data_ptr() == keys.data_ptr() and values.numel() == keys.numel()
    variant = _exchange_variant(keys, ops)
    fusable = keys.dtype == torch.int64 and (same or values.element_size() in (4, 8)) and hasattr(ops, "range_exchange")
    if variant == "fused" and fusable:
        ph.mark("partition+exchange(fused pass)")
        rk, rv = ops.range_exchange(keys, None if same else values, splitters, group)
        if same:
            rv = rk
    elif variant in ("fused", "staged", "plain"):
        if variant == "fused":
            variant = "staged"
        ph.mark("partition+exchange(p2p)")
        if same:
            rk = ops.partition_exchange(keys, keys, 0, splitters, group, variant=variant)
            rv = rk
        else:
            rk, rv = ops.partition_exchange([keys, values], keys, 0, splitters, group, variant=variant)
    else:
        ph.mark("partition")
        cols, offsets = ops.partition([keys] if same else [keys, values], keys, 0, splitters, world)
        ph.mark("exchange")
        recv = _exchange_cols(cols, offsets, ops, group)
        rk = recv[0]
        rv = rk if same else recv[1]
        del cols
    ph.mark("local_sort")
    out = ops.sort_by_key(rv, rk)
    ph.done("sort_by_key_sharded")
    return out


def inner_join_sharded(left_keys: torch.Tensor, right_keys: torch.Tensor, ops=None, group=None):
    """Inner join of two row-sharded int64 key columns -> (left_global_row, right_global_row) pairs found on this rank."""
    ops = ops or CudaOps()
    world, rank = dist.get_world_size(group), dist.get_rank(group)

    def global_ids(t):
        counts = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
        allc = [torch.empty_like(counts) for _ in range(world)]
        dist.all_gather(allc, counts, group=group)
        base = int(sum(int(c.item()) for c in allc[:rank]))
        return torch.arange(base, base + t.numel(), dtype=torch.int64, device=t.device)

    def shuffle(keys, slot_base):
        gid = global_ids(keys)
        if world == 1:
            return keys, gid
        variant = _exchange_variant(keys, ops)
        if variant == "fused" and keys.dtype == torch.int64 and hasattr(ops, "range_exchange"):
            # hash partition fused into ONE one-sweep pass (no splitters: bucket = hash of the key) carrying the global row id
            ph.mark("partition+exchange(fused pass)")
            got = ops.range_exchange(keys, gid, None, group, slot_base=slot_base)
            return got[0], got[1]
        if variant == "fused":
            variant = "staged"
        if variant in ("staged", "plain"):
            ph.mark("partition+exchange(p2p)")
            got = ops.partition_exchange([keys, gid], keys, 1, None, group, slot_base=slot_base, variant=variant)
            return got[0], got[1]
        ph.mark("partition")
        cols, offsets = ops.partition([keys, gid], keys, 1, None, world)
        ph.mark("exchange")
        got = _exchange_cols(cols, offsets, ops, group, slot_base)  # each side has its own pair of exchange buffers
        return got[0], got[1]

    ph = _Phase()
    lk, lg = shuffle(left_keys, 0)
    rk, rg = shuffle(right_keys, 2)
    ph.mark("local_join")
    li, ri = ops.inner_join(lk, rk)
    ph.mark("row_ids")
    out = lg[li.long()], rg[ri.long()]
    ph.done("inner_join_sharded")
    return out


def reduce_sharded(col: torch.Tensor, kind: str = "sum", ops=None, group=None):
    """local reduce + all_reduce of one scalar per rank (kind: sum | min | max)."""
    ops = ops or CudaOps()
    local = ops.reduce(col, kind)  # 0-d tensor on the column's device
    op = {"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}[kind]
    out = local.clone().reshape(1)
    if dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=op, group=group)
    return out[0]


def groupby_sum_count_sharded(keys: torch.Tensor, values: torch.Tensor, ops=None, group=None):
    """SURVEY 8e groupby: pre-aggregate locally (<= G rows), all_gather the partials (G is small), merge locally.
    Returns (group keys, sums, counts), identical on every rank, in arbitrary group order."""
    ops = ops or CudaOps()
    world = dist.get_world_size(group)
    k, s, c = ops.groupby_sum_count(keys, values)
    if world == 1:
        return k, s, c
    n = torch.tensor([k.numel()], dtype=torch.int64, device=k.device)
    sizes = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x.item()) for x in sizes]
    m = max(sizes + [1])

    def gather_var(t):
        pad = torch.zeros(m, dtype=t.dtype, device=t.device)
        pad[: t.numel()] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        return torch.cat([p[:sz] for p, sz in zip(parts, sizes)])

    ak, as_, ac = gather_var(k), gather_var(s), gather_var(c.to(torch.int64))
    return ops.groupby_merge(ak, as_, ac)
