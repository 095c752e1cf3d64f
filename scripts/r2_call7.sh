#!/usr/bin/env bash
# Round 2, GPU call 7 (1 GPU, short): ballot-based segment fix-up; fused hash partition; which barrier forms synccheck accepts.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2c7
mkdir -p "$O"
step() { local name=$1 to=$2; shift 2; local t0=$SECONDS; timeout "$to" "$@" > "$O/$name.log" 2>&1; echo "$name exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"; }
: > "$O/summary.txt"
step fixup_probe 200 python scripts/fixup_probe.py
step tests_sort 300 python -m pytest tests/test_sort_gpu.py tests/test_sort_hybrid_gpu.py tests/test_zzzz_range_partition.py -q -m gpu -x
for v in 0 1 2 3; do
  step barrier_$v 60 compute-sanitizer --tool synccheck --error-exitcode 3 scripts/ubench/barrier_probe $v
  grep -h "variant\|ERROR SUMMARY\|Barrier error" "$O/barrier_$v.log" | sort | uniq -c | head -5 | tee -a "$O/summary.txt"
done
tail -4 "$O/fixup_probe.log"
cat "$O/summary.txt"
