#!/usr/bin/env bash
# Round 2, GPU call 4 (1 GPU): full GPU suite with durations (as the driver runs it), bench (default line), one-sweep ranking variants,
# compute-sanitizer on small inputs.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2c4
mkdir -p "$O"
step() { local name=$1 to=$2; shift 2; local t0=$SECONDS; timeout "$to" "$@" > "$O/$name.log" 2>&1; echo "$name exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"; }
: > "$O/summary.txt"
step tests_all 1100 python -m pytest tests -q -m gpu -x --durations=25
step bench 600 python bench.py --steps 10 --warmup 3
B2_SORT_CFG=10 step bench_cfg10 200 python bench.py --no-e2e --no-ops --steps 3 --cpu-rows 100000
B2_SORT_CFG=11 step bench_cfg11 200 python bench.py --no-e2e --no-ops --steps 3 --cpu-rows 100000
B2_SORT_CFG=13 step bench_cfg13 200 python bench.py --no-e2e --no-ops --steps 3 --cpu-rows 100000
step sanitize_memcheck 300 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_small.py
step sanitize_racecheck 400 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_small.py
B2_SORT_CFG=10 step sanitize_racecheck_safe 400 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_small.py
step sanitize_synccheck 300 compute-sanitizer --tool synccheck --error-exitcode 3 python scripts/sanitize_small.py
for f in sanitize_memcheck sanitize_racecheck sanitize_racecheck_safe sanitize_synccheck; do tail -c 6000 "$O/$f.log" > "$O/$f.tail.txt"; grep -c "Race reported\|ERROR SUMMARY\|Invalid\|hazard" "$O/$f.log" > "$O/$f.count.txt" 2>/dev/null; rm -f "$O/$f.log"; done
du -sh "$O"
cat "$O/summary.txt"
