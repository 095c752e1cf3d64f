"""Times sort_by_key and its segment fix-up on two key distributions at --rows rows: uniform 64-bit keys (0.23 rows per segment after
four passes at 1e9 rows) and keys with the top three bits cleared (what one of 8 ranks of the sharded sort receives: 1.9 rows per
segment)."""
import argparse
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import cudf_b200.pylibcudf as plc
from cudf_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000_000)
a = ap.parse_args()
n = a.rows
keys = torch.empty(n, dtype=torch.int64, device="cuda")
_lib.check(_lib.lib.b2_fill_splitmix64(C.c_void_p(keys.data_ptr()), n, 0x5EED0001, 0, 0, 0, _lib.stream_arg(None)))
for name, k in (("uniform", keys), ("top3_clear", (keys >> 3) & 0x1FFFFFFFFFFFFFFF)):
    t = plc.Table([plc.Column.from_torch(k)])
    for _ in range(2):
        o = plc.sorting.sort_by_key(t, t, [0], [])
        del o
    _lib.lib.b2_profile_reset()
    _lib.lib.b2_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        o = plc.sorting.sort_by_key(t, t, [0], [])
        del o
    e1.record()
    torch.cuda.synchronize()
    _lib.lib.b2_profile_enable(0)
    fx, fc = _lib.profile_get("segment_fix")
    os_, oc = _lib.profile_get_over("onesweep", 0.2)
    hs, hc = _lib.profile_get("histogram")
    print(f"{name}: sort_by_key {e0.elapsed_time(e1) / 3:.2f} ms; segment_fix {fx / max(fc, 1):.2f} ms; onesweep {os_ / max(oc, 1):.2f} ms x {oc / 3:.0f}; histogram {hs / 3:.2f} ms")
    del t, k
