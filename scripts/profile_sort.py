"""Small driver for ncu captures: one sort_by_key over --rows int64 keys (after one warm-up call)."""
import argparse
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import cudf_b200.pylibcudf as plc
from cudf_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1 << 27)
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
keys = torch.empty(a.rows, dtype=torch.int64, device="cuda")
_lib.check(_lib.lib.b2_fill_splitmix64(C.c_void_p(keys.data_ptr()), a.rows, 0x5EED0001, 0, 0, 0, _lib.stream_arg(None)))
t = plc.Table([plc.Column.from_torch(keys)])
for _ in range(1 + a.reps):
    out = plc.sorting.sort_by_key(t, t, [0], [])
    torch.cuda.synchronize()
print("done", out.columns()[0].size())
