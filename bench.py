#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json): rows/s of cudf::sort_by_key on a
1e9-row int64 column (configs[1]) per GPU, through the pylibcudf-named shim over the C ABI.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--rows R] [--impl ours|reference]

One "step" = one sort_by_key(values=T, keys=T) over a resident synthetic column (splitmix64 keys,
SURVEY §8d).  `value` is device-resident throughput; `e2e` includes the pinned-host -> device copy of
the keys and the device -> pinned-host copy of the sorted column inside the timed region.
`roofline` is measured live with CUDA events around the one-sweep pass launches (b2_profile_*).
`cpu_baseline` / `--impl reference` time pandas sort_values (configs[0], the reference's CPU-runnable
case) on a bounded 1e7-row sample of the same key stream on the host cores.
Inputs (8 GB per GPU) are far larger than the 126 MB L2, so no explicit L2 flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "sort_by_key_rows_per_s"
UNIT = "rows/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rows", type=int, default=1_000_000_000, help="rows per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-rows", type=int, default=10_000_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-alias", action="store_true", help="skip the secondary measurement of the opt-in aliased (keys-only) sort path")
    ap.add_argument("--extra", action="store_true", help="also time join / groupby / scan / reduce (extra JSON keys)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU baseline (pandas; the oracle's host-side twin of the workload)
# ------------------------------------------------------------------------------------------------
def cpu_sort_sample(cpu_rows: int, steps: int, warmup: int):
    import numpy as np
    import pandas as pd

    from oracle import datagen

    keys = datagen.fill(cpu_rows, datagen.SEED_KEYS, 0, 0)
    df = pd.DataFrame({"k": keys})
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = df.sort_values("k", kind="stable")
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    assert bool(np.all(np.diff(out["k"].to_numpy()) >= 0))
    best = min(times)
    # other host implementations of the same sort on the same sample (one run each; reported, not the baseline value)
    alt = {}
    try:
        t0 = time.perf_counter()
        np.sort(keys, kind="stable")
        alt["numpy_stable_sort_rows_per_s"] = cpu_rows / (time.perf_counter() - t0)
        import pyarrow as pa
        import pyarrow.compute as pc

        arr = pa.array(keys)
        t0 = time.perf_counter()
        pc.take(arr, pc.sort_indices(arr))
        alt["pyarrow_sort_indices_take_rows_per_s"] = cpu_rows / (time.perf_counter() - t0)
        alt["pyarrow_threads"] = pa.cpu_count()
    except Exception as ex:  # optional extras only
        alt["error"] = repr(ex)[:120]
    return {
        "value": cpu_rows / best,
        "unit": UNIT,
        "cores": 1,  # pandas/numpy sort is single-threaded
        "kind": "port",
        "sample": f"pandas {pd.__version__} DataFrame.sort_values(kind='stable') on {cpu_rows} int64 rows "
                  f"(same splitmix64 key stream), best of {steps}; host has {os.cpu_count()} logical cores",
        "ms": best * 1e3,
        "alternatives": alt,
    }, sum(times) / len(times)


def workload_config(n: int, world: int) -> dict:
    return {"workload": f"{n}-row single int64 column sort_by_key(values=T, keys=T), no nulls, ASCENDING, per GPU "
                        "(BASELINE.json configs[1])" + ("" if world == 1 else f"; sharded over {world} GPUs: sample-sort splitters, stable range "
                        "partition, bucket exchange (fused peer-memory scatter over NVLink at 2 ranks, NCCL all-to-all-v "
                        "above), local LSD sort (configs[4])"),
            "rows_per_gpu": n, "l2_flush": "inputs (8 GB/GPU) exceed the 126 MB L2; no explicit flush",
            "generator": "splitmix64(seed 0x5EED0001 + i)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 2))
    base, mean_s = cpu_sort_sample(args.cpu_rows, steps, warmup)
    line = {
        "impl": "reference",
        "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": dict(workload_config(args.rows, max(1, args.gpus)),
                       sample=f"each step = pandas DataFrame.sort_values(kind='stable') on a {args.cpu_rows}-row sample of the key "
                              "stream (BASELINE.json configs[0]: the reference's CPU-runnable case)", rows_per_step=args.cpu_rows),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                  "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        try:
            while not self._stop.is_set():
                line = p.stdout.readline()
                if not line:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        finally:
            p.terminate()

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx.append(float(s[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm_sorted = sorted(sm)
        # "under load": upper half of the samples (the sampler also sees the idle gaps between steps)
        load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__ as g

    g.build()
    import cudf_b200.pylibcudf as plc
    from cudf_b200 import _lib
    from oracle import datagen

    n = args.rows
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream()
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    # disjoint counter range per rank: rank r draws x_i for i in [r*n, (r+1)*n)
    _lib.check(_lib.lib.b2_fill_splitmix64(C.c_void_p(keys.data_ptr()), n, datagen.SEED_KEYS, rank * n, 0, 0, _lib.stream_arg(None)))
    torch.cuda.synchronize()

    if world > 1:
        from cudf_b200 import sharded

        def step():
            return sharded.sort_by_key_sharded(keys, keys)
    else:
        col = plc.Column.from_torch(keys)
        tbl = plc.Table([col])

        def step():
            return plc.sorting.sort_by_key(tbl, tbl, [plc.Order.ASCENDING], [])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step()
        del out
    barrier()

    # ---- timed region: device-resident ----
    _lib.lib.b2_profile_reset()
    _lib.lib.b2_profile_enable(1)
    launches0 = _lib.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            out = step()
            del out
        e1.record(stream)
        barrier()
    _lib.lib.b2_profile_enable(0)
    ms = e0.elapsed_time(e1)
    launches = _lib.kernel_launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * n / (ms_step / 1e3)

    # ---- live roofline of the dominant kernel (one-sweep pass) ----
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if peaks_path.exists():
        try:
            peak = float(json.loads(peaks_path.read_text())["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    os_ms, os_cnt = _lib.profile_get("onesweep")
    hist_ms, hist_cnt = _lib.profile_get("histogram")
    ga_ms, ga_cnt = _lib.profile_get("gather")
    fix_ms, fix_cnt = _lib.profile_get("segment_fix")
    # algorithmic bytes of THIS implementation's 8 passes over (int64 key, int32 row id):
    # pass 1: 8 read + 12 write; passes 2-7: 12 + 12; pass 8: 12 read + 4 write (row ids only) = 180 B/row
    rows_local = n  # per rank; at N > 1 the received shard differs from n by < 1 % (sample-sort splitters)
    # opt-in sort paths (README "Environment switches") move different bytes per pass; the default is the row-id path
    if os.environ.get("B2_SORT_ALIAS", "0") not in ("", "0"):
        sort_path, bytes_8_passes, kernel_name = "aliased keys-only radix (B2_SORT_ALIAS)", 128.0, "onesweep_kernel<uint64, keys only>"
    elif os.environ.get("B2_SORT_CARRY", "0") not in ("", "0"):
        sort_path, bytes_8_passes, kernel_name = "payload-carrying radix (B2_SORT_CARRY)", 248.0, "onesweep_kernel<uint64,(key,8-byte payload)>"
    else:
        sort_path, bytes_8_passes, kernel_name = "row ids + gather (default)", 180.0, "onesweep_kernel<uint64,(key,row id)>"
    default_path = bytes_8_passes == 180.0
    roofline = None
    if os_cnt and rows_local:
        per_launch_bytes = bytes_8_passes * rows_local / 8.0
        # the n-row sort runs 8 passes per step (uniform 64-bit keys: no trivial digit); at N > 1 the splitter sample sort
        # adds a few microsecond-scale launches per step, whose time stays in the numerator and is negligible
        big_launches = 8 * args.steps
        avg_ms = os_ms / big_launches
        achieved = per_launch_bytes / (avg_ms / 1e3) / 1e9
        roofline = {
            "bound": "hbm", "kernel": kernel_name, "sort_path": sort_path, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the ncu --set full capture at 2^27 rows
            # (profiles/r1_onesweep_ncu_e.txt: 1.612 + 1.582 GB per launch = 23.8 B/row), scaled to this launch size
            "traffic": 23.8 * rows_local if default_path else None,
            "traffic_source": "ncu capture at 2^27 rows, per-row figure scaled" if default_path else "no ncu capture of this path yet",
            "peak_source": peak_src, "rank": rank,
            "algorithmic_bytes_per_launch": per_launch_bytes, "avg_launch_ms": avg_ms, "launches": big_launches, "launches_incl_sample_sort": os_cnt,
            "kernel_share_of_step": os_ms / ms_total,
            "whole_op": {"algorithmic_bytes_per_row_contract": 216, "achieved_GBps_contract": 216.0 * rows_local / (ms_step / 1e3) / 1e9,
                         "frac_contract": 216.0 * rows_local / (ms_step / 1e3) / 1e9 / peak},
            "other_kernels_ms_per_step": {"histogram": hist_ms / args.steps, "gather": ga_ms / args.steps, "segment_fix": fix_ms / args.steps,
                                          "onesweep_total": os_ms / args.steps},
        }

    # ---- e2e: pinned host -> device -> sort_by_key -> pinned host ----
    e2e = None
    if not args.no_e2e:
        try:
            cap = n if world == 1 else int(n * 1.1) + 1024  # a rank's shard of the sharded result is ~n rows
            alloc_err = None
            try:
                h_in = torch.empty(n, dtype=torch.int64, pin_memory=True)
                h_out = torch.empty(cap, dtype=torch.int64, pin_memory=True)
                d_in = torch.empty(n, dtype=torch.int64, device=dev)
            except Exception as ex:  # e.g. not enough pinnable host memory on this rank
                alloc_err = ex
            ok = torch.tensor([0 if alloc_err else 1], dtype=torch.int32, device=dev)
            if world > 1:  # every rank must take the same branch, or the collectives below would hang
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                raise RuntimeError(f"e2e buffers could not be allocated on every rank: {alloc_err!r}")
            h_in.copy_(keys)
            torch.cuda.synchronize()

            if world == 1:
                def e2e_step():
                    d_in.copy_(h_in, non_blocking=True)
                    c = plc.Column.from_torch(d_in)
                    o = plc.sorting.sort_by_key(plc.Table([c]), plc.Table([c]), [plc.Order.ASCENDING], [])
                    h_out.copy_(o.columns()[0].to_torch(), non_blocking=True)
                    return o
            else:
                def e2e_step():
                    d_in.copy_(h_in, non_blocking=True)
                    o = sharded.sort_by_key_sharded(d_in, d_in)
                    m = min(o.numel(), cap)
                    h_out[:m].copy_(o[:m], non_blocking=True)
                    return o

            o = e2e_step()
            torch.cuda.synchronize()
            del o
            k = max(1, min(args.steps, 3))
            barrier()
            e0.record(stream)
            for _ in range(k):
                o = e2e_step()
                del o
            e1.record(stream)
            barrier()
            te = torch.tensor([e0.elapsed_time(e1) / k], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            ems = float(te.item())
            assert bool((h_out[1:1000001] >= h_out[:1000000]).all())
            e2e = {"value": world * n / (ems / 1e3), "unit": UNIT, "h2d_bytes_per_step": 8 * n * world,
                   "d2h_bytes_per_step": 8 * n * world, "ms_per_step": ems}
            del h_in, h_out, d_in
        except Exception as ex:  # e.g. not enough pinnable host memory
            e2e = {"value": None, "unit": UNIT, "error": repr(ex)[:200]}

    # ---- secondary, N = 1 only: the opt-in aliased path (sort_by_key(T, T) of one null-free integer column routed to the
    # keys-only radix, README "Environment switches"). Reported next to the headline, never as the headline; its output is
    # compared with the default path's output first.
    alias = None
    if world == 1 and not args.no_alias and os.environ.get("B2_SORT_ALIAS", "0") in ("", "0"):
        try:
            ref_out = step().columns()[0].to_torch()
            os.environ["B2_SORT_ALIAS"] = "1"
            try:
                got = step().columns()[0].to_torch()
                same = bool(torch.equal(got, ref_out))
                del got, ref_out
                for _ in range(2):
                    o = step()
                    del o
                torch.cuda.synchronize()
                k = max(1, min(args.steps, 3))
                e0.record(stream)
                for _ in range(k):
                    o = step()
                    del o
                e1.record(stream)
                torch.cuda.synchronize()
                ams = e0.elapsed_time(e1) / k
                alias = {"path": "B2_SORT_ALIAS=1: keys-only radix, no row ids, no gather (opt-in, not the headline)", "ms_per_step": ams,
                         "rows_per_s": n / (ams / 1e3), "output_equals_default_path": same,
                         "algorithmic_bytes_per_row": 136, "achieved_GBps": 136.0 * n / (ams / 1e3) / 1e9}
            finally:
                os.environ.pop("B2_SORT_ALIAS", None)
        except Exception as ex:
            alias = {"error": repr(ex)[:200]}

    extra = None
    if args.extra and world == 1:
        import bench_extra

        del keys
        extra = bench_extra.run(plc, _lib, n, peak)

    if rank == 0:
        cpu_base, _ = cpu_sort_sample(args.cpu_rows, 3, 1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic",
            "config": workload_config(n, world),
            "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        if alias:
            line["aliased_keys_only_path"] = alias
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
