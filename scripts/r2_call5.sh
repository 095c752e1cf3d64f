#!/usr/bin/env bash
# Round 2, GPU call 5 (2 GPUs): new one-sweep defaults (race-free ranking + ATOMS.ADD offsets, explicit-count barriers) under the
# sanitizer and in the 1-GPU bench, then the fused range-partition exchange against the staged scatter at 2 GPUs.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2c5
mkdir -p "$O"
step() { local name=$1 to=$2; shift 2; local t0=$SECONDS; timeout "$to" "$@" > "$O/$name.log" 2>&1; echo "$name exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"; }
: > "$O/summary.txt"
export CUDA_VISIBLE_DEVICES_SAVE=${CUDA_VISIBLE_DEVICES:-}
step tests_sort 300 python -m pytest tests/test_sort_gpu.py tests/test_sort_hybrid_gpu.py tests/test_zzzz_range_partition.py -q -m gpu -x
step bench1 300 python bench.py --no-e2e --no-ops --steps 5 --cpu-rows 100000
step sanitize_racecheck 300 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_small.py
step sanitize_synccheck 300 compute-sanitizer --tool synccheck --error-exitcode 3 python scripts/sanitize_small.py
for f in sanitize_racecheck sanitize_synccheck; do tail -c 3000 "$O/$f.log" > "$O/$f.tail.txt"; rm -f "$O/$f.log"; done
bash scripts/r2_multi.sh 2 fused staged
cp gpurun_out/r2multi2/summary.txt "$O/multi2_summary.txt" 2>/dev/null
cat "$O/summary.txt"
