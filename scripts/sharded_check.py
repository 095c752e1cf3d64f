"""torchrun --nproc-per-node N scripts/sharded_check.py [--rows R]: correctness of the sharded sort / join on N GPUs."""
import argparse
import ctypes as C
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from cudf_b200 import _lib, sharded

n = a.rows
keys = torch.empty(n, dtype=torch.int64, device="cuda")
_lib.check(_lib.lib.b2_fill_splitmix64(C.c_void_p(keys.data_ptr()), n, 0x5EED0001, rank * n, 0, 0, _lib.stream_arg(None)))
out = sharded.sort_by_key_sharded(keys, keys)
torch.cuda.synchronize()
# properties: locally sorted; boundaries ordered across ranks; global multiset preserved (sum / xor / count)
assert bool((out[1:] >= out[:-1]).all())
lo = out[:1] if out.numel() else keys.new_full((1,), 2**62)
hi = out[-1:] if out.numel() else keys.new_full((1,), -2**62)
los = [torch.empty_like(lo) for _ in range(world)]; his = [torch.empty_like(hi) for _ in range(world)]
dist.all_gather(los, lo); dist.all_gather(his, hi)
for r in range(world - 1):
    assert int(his[r]) <= int(los[r + 1]), (r, int(his[r]), int(los[r + 1]))
stats_in = torch.stack([keys.sum(), torch.tensor(keys.numel(), device="cuda"), (keys ^ (keys >> 7)).sum()])
stats_out = torch.stack([out.sum(), torch.tensor(out.numel(), device="cuda"), (out ^ (out >> 7)).sum()])
dist.all_reduce(stats_in); dist.all_reduce(stats_out)
assert torch.equal(stats_in, stats_out), (stats_in, stats_out)
cnt = torch.tensor([out.numel()], device="cuda"); cnts = [torch.empty_like(cnt) for _ in range(world)]; dist.all_gather(cnts, cnt)
# payload column distinct from the keys: values = 3 * key + 1 (wrapping) must arrive in key order
vals = keys * 3 + 1
outv = sharded.sort_by_key_sharded(vals, keys)
torch.cuda.synchronize()
assert outv.numel() == out.numel() and bool((outv == out * 3 + 1).all())
v32 = (keys & 0x7FFFFFFF).to(torch.int32)
outv32 = sharded.sort_by_key_sharded(v32, keys)
torch.cuda.synchronize()
assert bool((outv32 == (out & 0x7FFFFFFF).to(torch.int32)).all())
# join: right = fresh keys, left = 10 % copies of right rows of ANY rank via the shared generator
m = max(1000, n // 10)
rk = torch.empty(m, dtype=torch.int64, device="cuda")
_lib.check(_lib.lib.b2_fill_splitmix64(C.c_void_p(rk.data_ptr()), m, 0x5EED0002, rank * m, 0, 0, _lib.stream_arg(None)))
lk = torch.empty(m, dtype=torch.int64, device="cuda")
# left rank r copies the keys of right rank (r+1) % world for even rows, fresh keys otherwise
_lib.check(_lib.lib.b2_fill_splitmix64(C.c_void_p(lk.data_ptr()), m, 0x5EED0002, ((rank + 1) % world) * m, 0, 0, _lib.stream_arg(None)))
fresh = torch.empty(m, dtype=torch.int64, device="cuda")
_lib.check(_lib.lib.b2_fill_splitmix64(C.c_void_p(fresh.data_ptr()), m, 0x5EED0009, (1 << 45) + rank * m, 0, 0, _lib.stream_arg(None)))
even = (torch.arange(m, device="cuda") % 2) == 0
lk = torch.where(even, lk, fresh)
jl, jr = sharded.inner_join_sharded(lk, rk)
tot = torch.tensor([jl.numel()], device="cuda"); dist.all_reduce(tot)
exp = world * int(even.sum())
assert int(tot) == exp, (int(tot), exp)
# each pair: left global id g -> rank g // m, row g % m, must be an even row; right id = ((lrank+1)%world)*m + row
lrank, lrow = jl // m, jl % m
assert bool((lrow % 2 == 0).all()) and bool((jr == ((lrank + 1) % world) * m + lrow).all())
if rank == 0:
    print(f"SHARDED_OK world={world} rows/rank={n} shard sizes={[int(c) for c in cnts]} join pairs={int(tot)}")
dist.destroy_process_group()
