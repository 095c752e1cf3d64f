#!/usr/bin/env bash
# Multi-GPU call: correctness (scripts/sharded_check.py) and timing (bench.py) of each bucket-exchange variant.
#   gpurun --gpus N --timeout 1500 -- 'bash scripts/r2_multi.sh N [variants...]'
# variants: peer (partition + contiguous peer copies) | staged (fused staged scatter) | p2p (fused plain scatter) | nccl
set -u
cd "$(dirname "$0")/.."
N=${1:-2}; shift || true
VARIANTS=${*:-"peer staged p2p nccl"}
O=gpurun_out/r2multi$N
mkdir -p "$O"
: > "$O/summary.txt"
run() {  # run <name> <timeout> <env...> -- <cmd...>
  local name=$1 to=$2; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  local t0=$SECONDS
  env "${envs[@]}" timeout "$to" "$@" > "$O/$name.log" 2>&1
  echo "$name exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > "$O/topo.txt" 2>&1
for v in $VARIANTS; do
  case $v in
    peer)   E="B2_SHARD_P2P=copy B2_SHARD_XCHG=peer" ;;
    staged) E="B2_SHARD_P2P=staged" ;;
    fused)  E="X=1" ;;
    copy)   E="B2_SHARD_P2P=copy" ;;
    p2p)    E="B2_SHARD_P2P=1" ;;
    nccl)   E="B2_SHARD_P2P=copy B2_SHARD_XCHG=nccl" ;;
    default) E="X=1" ;;
  esac
  run check_$v 150 $E -- $TR scripts/sharded_check.py --rows 20000000
  run bench_$v 240 $E -- $TR bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --no-join --cpu-rows 100000
done
# the default path with everything (e2e + sharded join) once
run bench_default 420 X=1 -- $TR bench.py --gpus $N --steps 5 --warmup 3 --cpu-rows 1000000
grep -h '"metric"' "$O"/bench_*.log | python -c "
import sys, json
for line in sys.stdin:
    try:
        d = json.loads(line)
    except Exception:
        continue
    print(d['n_gpus'], round(d['ms_per_step'], 2), 'ms', round(d['value'] / 1e9, 2), 'Grows/s', d.get('phases_ms'), (d.get('sharded_inner_join') or {}).get('ms_per_step'), (d.get('e2e') or {}).get('ms_per_step'))
" | tee -a "$O/summary.txt"
cat "$O/summary.txt"
