#!/usr/bin/env bash
# Round 2, GPU call 9 (1 GPU): fix-up reverted to the branch-free walk; rj2 join kernel (B2_JOIN_KERNEL=2) and histogram-free groupby
# partition (B2_GROUPBY_EST=1) against the defaults, each in fresh processes with per-iteration times; default bench line with the
# cached memory released between operation groups.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2c9
mkdir -p "$O"
step() { local name=$1 to=$2; shift 2; local t0=$SECONDS; timeout "$to" "$@" > "$O/$name.log" 2>&1; echo "$name exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"; }
: > "$O/summary.txt"
step tests_sort 300 python -m pytest tests/test_sort_gpu.py tests/test_sort_hybrid_gpu.py tests/test_groupby_partitioned_gpu.py -q -m gpu -x
B2_JOIN_KERNEL=2 step tests_join_k2 400 python -m pytest tests/test_parity_gpu.py tests/test_zz_experimental_gpu.py tests/test_zzzz_match_context.py tests/test_zz_full_size_gpu.py -q -m gpu -x -k "join or radix or match"
B2_JOIN_KERNEL=1 step probe_join_k1 200 python scripts/ops_probe.py --op join
B2_JOIN_KERNEL=2 step probe_join_k2 200 python scripts/ops_probe.py --op join
B2_GROUPBY_EST=0 step probe_gb_est0 200 python scripts/ops_probe.py --op groupby
B2_GROUPBY_EST=1 step probe_gb_est1 200 python scripts/ops_probe.py --op groupby
B2_GROUPBY_EST=1 B2_GROUPBY_CHUNK=524288 step probe_gb_est1_c19 200 python scripts/ops_probe.py --op groupby
B2_GROUPBY_EST=1 B2_GROUPBY_CHUNK=1048576 step probe_gb_est1_c20 200 python scripts/ops_probe.py --op groupby
step fixup_probe 200 python scripts/fixup_probe.py
step bench 500 python bench.py --steps 8 --warmup 3
for f in probe_join_k1 probe_join_k2 probe_gb_est0 probe_gb_est1 probe_gb_est1_c19 probe_gb_est1_c20; do tail -1 "$O/$f.log" | cut -c1-900; done
tail -2 "$O/fixup_probe.log"
tail -3 "$O/tests_sort.log"; tail -3 "$O/tests_join_k2.log"
grep -h '"metric"' "$O/bench.log" | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print(d['ms_per_step'], d['e2e'].get('ms_per_step'))
for k, v in d['ops'].items(): print(k, round(v.get('ms', -1), 2), v.get('phases_ms'))
"
cat "$O/summary.txt"
