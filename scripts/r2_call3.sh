#!/usr/bin/env bash
# Round 2, GPU call 3 (1 GPU): fused radix join kernel, scan with 2 CTAs/SM, plan read-back, two-phase histogram; bench with ops; ncu of
# the partitioned groupby, the join kernel and the scan.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2c3
mkdir -p "$O"
step() { local name=$1 to=$2; shift 2; local t0=$SECONDS; timeout "$to" "$@" > "$O/$name.log" 2>&1; echo "$name exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"; }
: > "$O/summary.txt"
B2_RUN_EXPERIMENTAL=1 step tests_join 600 python -m pytest tests/test_zz_experimental_gpu.py tests/test_parity_gpu.py tests/test_zzzz_match_context.py -q -m gpu -x -k "join or radix or match or partitioned"
step tests_join_large 600 python -m pytest tests/test_zz_full_size_gpu.py -q -m gpu -x -k "inner_join"
step tests_new 600 python -m pytest tests/test_sort_hybrid_gpu.py tests/test_groupby_partitioned_gpu.py tests/test_zzzz_partitioning.py tests/test_zzzz_pack.py tests/test_zzz_cpp_api.py -q -m gpu -x
step tests_scan 300 python -m pytest tests/test_parity_gpu.py tests/test_golden_ops.py -q -m gpu -x -k "scan or reduce"
step bench 900 python bench.py --steps 5
R=134217728
cap() {  # cap <name> <kernel regex> <skip> <count> <env...> -- <op>
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  local t0=$SECONDS
  env "${envs[@]}" timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s "$skip" -c "$cnt" -f -o "$O/$name" \
    python scripts/profile_ops.py --op "$1" --rows $R > "$O/$name.log" 2>&1
  local rc=$?
  if [ -f "$O/$name.ncu-rep" ]; then
    ncu -i "$O/$name.ncu-rep" --page raw --csv > "$O/${name}_raw.csv" 2>/dev/null
    ncu -i "$O/$name.ncu-rep" --page source --csv > "$O/${name}_src.csv" 2>/dev/null
    rm -f "$O/$name.ncu-rep"
  fi
  echo "ncu_$name exit=$rc secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"
}
cap groupby "pgb_agg_kernel" 1 1 X=1 -- groupby
cap rjoin "rj_join_kernel" 1 1 X=1 -- inner_join
cap scan "scan_kernel" 1 1 X=1 -- scan
step ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$O/bench_launches.csv" \
  python bench.py --steps 2 --warmup 1 --no-e2e --cpu-rows 100000
du -sh "$O"
cat "$O/summary.txt"
