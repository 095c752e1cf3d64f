#!/usr/bin/env bash
# Multi-GPU confirmation of the default path only: correctness (scripts/sharded_check.py), then the default bench line without the
# host round trip (sort + sharded join with their phases).   gpurun --gpus N -- 'bash scripts/r2_multi_final.sh N'
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out/r2final$N
mkdir -p "$O"
: > "$O/summary.txt"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
t0=$SECONDS; timeout 150 $TR scripts/sharded_check.py --rows 20000000 > "$O/check.log" 2>&1; echo "check exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"
grep -h SHARDED_OK "$O/check.log" | tee -a "$O/summary.txt"
t0=$SECONDS; timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --cpu-rows 100000 > "$O/bench.log" 2>&1; echo "bench exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"
grep -h '"metric"' "$O/bench.log" | python -c "
import sys, json
for line in sys.stdin:
    d = json.loads(line)
    j = d.get('sharded_inner_join') or {}
    print(d['n_gpus'], round(d['ms_per_step'], 2), 'ms', round(d['value'] / 1e9, 2), 'Grows/s', d.get('phases_ms'), 'join', j.get('ms_per_step'), j.get('phases_ms'), j.get('row_ids_consistent'))
" | tee -a "$O/summary.txt"
tail -5 "$O/check.log" "$O/bench.log" | cut -c1-400
