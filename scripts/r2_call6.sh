#!/usr/bin/env bash
# Round 2, GPU call 6 (1 GPU, short): segment fix-up fast path on uniform and range-partitioned keys; synccheck with the named barrier.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2c6
mkdir -p "$O"
step() { local name=$1 to=$2; shift 2; local t0=$SECONDS; timeout "$to" "$@" > "$O/$name.log" 2>&1; echo "$name exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"; }
: > "$O/summary.txt"
step fixup_probe 200 python scripts/fixup_probe.py
step tests_sort 300 python -m pytest tests/test_sort_gpu.py tests/test_sort_hybrid_gpu.py -q -m gpu -x
step sanitize_synccheck 300 compute-sanitizer --tool synccheck --error-exitcode 3 python scripts/sanitize_small.py
tail -c 3000 "$O/sanitize_synccheck.log" > "$O/sanitize_synccheck.tail.txt"; rm -f "$O/sanitize_synccheck.log"
cat "$O/fixup_probe.log" | tail -4
cat "$O/summary.txt"
