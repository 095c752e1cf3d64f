#!/usr/bin/env bash
# Round 2, first GPU call: hardware verdict of the opt-in paths + evidence (bench with --extra, ncu launch list, full captures).
#   gpurun --timeout 1800 -- 'bash scripts/r2_call1.sh'
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2c1
mkdir -p "$O"
step() { local name=$1 to=$2; shift 2; local t0=$SECONDS; timeout "$to" "$@" > "$O/$name.log" 2>&1; echo "$name exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"; }
: > "$O/summary.txt"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > "$O/gpu.txt" 2>&1
B2_RUN_EXPERIMENTAL=1 step tests_experimental 900 python -m pytest tests/test_zz_experimental_gpu.py -q -m gpu -rxX
step tests_join 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "hash_join_object"
step bench 600 python bench.py --no-e2e
step bench_extra 900 python bench.py --extra --no-e2e --no-alias --steps 3
B2_SORT_CARRY=1 step bench_carry 400 python bench.py --no-e2e --no-alias --steps 3
B2_SORT_CFG=10 step bench_cfg10 400 python bench.py --no-e2e --no-alias --steps 3 --cpu-rows 100000
B2_SORT_CFG=11 step bench_cfg11 400 python bench.py --no-e2e --no-alias --steps 3 --cpu-rows 100000
step ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$O/bench_launches.csv" \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-alias --cpu-rows 100000
R=134217728
step ncu_sort 600 ncu --set full --clock-control none --import-source on -k "regex:onesweep|gather_kernel|histogram" -s 10 -c 10 -f -o "$O/sort_full" \
  python scripts/profile_ops.py --op sort_by_key --rows $R
B2_SORT_CARRY=1 step ncu_carry 600 ncu --set full --clock-control none --import-source on -k "regex:onesweep" -s 9 -c 2 -f -o "$O/carry_full" \
  python scripts/profile_ops.py --op sort_by_key_payload --rows $R
step ncu_join 600 ncu --set full --clock-control none --import-source on -k "regex:build_kernel|count_kernel|retrieve_kernel" -s 3 -c 3 -f -o "$O/join_full" \
  python scripts/profile_ops.py --op inner_join --rows $R
B2_JOIN_RADIX_ROWS=1 step ncu_rjoin_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$O/rjoin_launches.csv" \
  python scripts/profile_ops.py --op inner_join --rows $R
step ncu_groupby 600 ncu --set full --clock-control none --import-source on -k "regex:groupby_kernel" -s 1 -c 1 -f -o "$O/groupby_full" \
  python scripts/profile_ops.py --op groupby --rows $R
step ncu_scan 600 ncu --set full --clock-control none --import-source on -k "regex:scan_kernel" -s 1 -c 1 -f -o "$O/scan_full" \
  python scripts/profile_ops.py --op scan --rows $R
cat "$O/summary.txt"
