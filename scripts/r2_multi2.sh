#!/usr/bin/env bash
# Multi-GPU confirmation of the defaults (fused range-partition pass for the sort, fused hash-partition pass for the join's shuffle)
# next to the staged scatter:   gpurun --gpus N -- 'bash scripts/r2_multi2.sh N'
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out/r2m$N
mkdir -p "$O"
: > "$O/summary.txt"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { local name=$1 to=$2; shift 2; local t0=$SECONDS; timeout "$to" "$@" > "$O/$name.log" 2>&1; echo "$name exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"; }
run check 150 $TR scripts/sharded_check.py --rows 20000000
grep -h SHARDED_OK "$O/check.log" | tee -a "$O/summary.txt"
run bench_default 300 $TR bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --cpu-rows 100000
B2_SHARD_P2P=staged run bench_staged 300 $TR bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --no-join --cpu-rows 100000
for f in bench_default bench_staged; do
grep -h '"metric"' "$O/$f.log" | python -c "
import sys, json
for line in sys.stdin:
    d = json.loads(line)
    j = d.get('sharded_inner_join') or {}
    print('$f', d['n_gpus'], round(d['ms_per_step'], 2), 'ms', round(d['value'] / 1e9, 2), 'Grows/s', d.get('phases_ms'), 'join', j.get('ms_per_step'), j.get('phases_ms'), j.get('row_ids_consistent'))
" | tee -a "$O/summary.txt"
done
tail -5 "$O/check.log" "$O/bench_default.log" | cut -c1-300
