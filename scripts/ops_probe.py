"""Per-iteration timing of one BASELINE operation in a fresh process (the kernel / path switches are read once per process):
   python scripts/ops_probe.py --op join|groupby [--rows N] [--iters K]
Prints one line per iteration (CUDA events around the call) and the profile-hook phases averaged over the iterations after the first
two, so that allocator warm-up and steady state can be told apart. Inputs as in bench_extra.py (BASELINE configs[2] / configs[3])."""
import argparse
import ctypes as C
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import cudf_b200.pylibcudf as plc
from cudf_b200 import _lib
from bench_extra import _fill

ap = argparse.ArgumentParser()
ap.add_argument("--op", required=True, choices=["join", "groupby"])
ap.add_argument("--rows", type=int, default=1_000_000_000)
ap.add_argument("--iters", type=int, default=6)
a = ap.parse_args()
n, dev = a.rows, "cuda"
agg = plc.aggregation
if a.op == "groupby":
    G = 1_000_000
    k = _fill(_lib, torch.empty(n, dtype=torch.int64, device=dev), n, 9, kind=2, modulus=G)
    f = _fill(_lib, torch.empty(n, dtype=torch.float64, device=dev), n, 8, kind=1)
    v2 = _fill(_lib, torch.empty(n, dtype=torch.int32, device=dev), n, 10, kind=3)
    gb = plc.groupby.GroupBy(plc.Table([plc.Column.from_torch(k)]))
    reqs = [plc.groupby.GroupByRequest(plc.Column.from_torch(f), [agg.sum()]), plc.groupby.GroupByRequest(plc.Column.from_torch(v2), [agg.count()])]
    fn = lambda: gb.aggregate(reqs)
    names = ("groupby_partition", "groupby_aggregate", "histogram", "onesweep")
    size = lambda o: o[0].num_rows()
else:
    rk = _fill(_lib, torch.empty(n, dtype=torch.int64, device=dev), n, 1)
    lk = _fill(_lib, torch.empty(n, dtype=torch.int64, device=dev), n, 6)
    u = _fill(_lib, torch.empty(n, dtype=torch.float64, device=dev), n, 5, kind=1)
    sel = _fill(_lib, torch.empty(n, dtype=torch.int64, device=dev), n, 4, kind=2, modulus=n)
    hit = u < 0.10
    del u
    lk[hit] = rk[sel[hit]]
    del sel, hit
    L, R = plc.Table([plc.Column.from_torch(lk)]), plc.Table([plc.Column.from_torch(rk)])
    fn = lambda: plc.join.inner_join(L, R, plc.NullEquality.EQUAL)
    names = ("rjoin_partition", "rjoin_join", "histogram", "onesweep")
    size = lambda o: o[0].size()
torch.cuda.synchronize()
torch.cuda.empty_cache()
times = []
for it in range(a.iters):
    if it == 2:
        _lib.lib.b2_profile_reset()
        _lib.lib.b2_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    o = fn()
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
    m = size(o)
    del o
_lib.lib.b2_profile_enable(0)
ph = {}
for k_ in names:
    t, c = _lib.profile_get(k_)
    if c:
        ph[k_] = {"ms_per_launch": t / c, "launches_per_iter": c / max(a.iters - 2, 1)}
env = {k_: v for k_, v in os.environ.items() if k_.startswith("B2_")}
print(json.dumps({"op": a.op, "rows": n, "env": env, "result_rows": int(m), "ms_per_iter": [round(t, 2) for t in times],
                  "steady_ms": round(sum(times[2:]) / max(len(times) - 2, 1), 2), "phases": ph}))
