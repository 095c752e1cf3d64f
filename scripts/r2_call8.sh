#!/usr/bin/env bash
# Round 2, GPU call 8 (1 GPU): full GPU suite on the current tree (ballot fix-up, fused hash partition), fix-up probe, default bench line.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2c8
mkdir -p "$O"
step() { local name=$1 to=$2; shift 2; local t0=$SECONDS; timeout "$to" "$@" > "$O/$name.log" 2>&1; echo "$name exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"; }
: > "$O/summary.txt"
step tests_all 900 python -m pytest tests -q -m gpu -x --durations=10
step fixup_probe 200 python scripts/fixup_probe.py
step bench 500 python bench.py --steps 8 --warmup 3
tail -4 "$O/fixup_probe.log"
tail -15 "$O/tests_all.log"
grep -h '"metric"' "$O/bench.log" | cut -c1-1500
cat "$O/summary.txt"
