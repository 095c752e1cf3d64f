#!/usr/bin/env bash
# Round 2, GPU call 2: hybrid + carry sort (new defaults) vs the round-1 path, opt-in paths, bench --extra, ncu evidence.
# ncu reports are exported to CSV on the box and deleted (gpurun_out/ is capped at 64 MiB).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2c2
mkdir -p "$O"
step() { local name=$1 to=$2; shift 2; local t0=$SECONDS; timeout "$to" "$@" > "$O/$name.log" 2>&1; echo "$name exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"; }
: > "$O/summary.txt"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > "$O/gpu.txt" 2>&1
step tests_hybrid 600 python -m pytest tests/test_sort_hybrid_gpu.py -q -m gpu -x
B2_RUN_EXPERIMENTAL=1 step tests_experimental 900 python -m pytest tests/test_zz_experimental_gpu.py -q -m gpu -rxX
step tests_all 1200 python -m pytest tests -q -m gpu -x --ignore tests/test_zz_experimental_gpu.py --ignore tests/test_sort_hybrid_gpu.py
step bench 600 python bench.py --no-e2e --no-ops
B2_SORT_HYBRID=0 B2_SORT_CARRY=0 step bench_r1path 400 python bench.py --no-e2e --no-ops --steps 3 --cpu-rows 100000
B2_SORT_HYBRID=0 B2_SORT_CARRY=1 step bench_carry_only 400 python bench.py --no-e2e --no-ops --steps 3 --cpu-rows 100000
B2_SORT_HYBRID=1 B2_SORT_CARRY=0 step bench_hybrid_only 400 python bench.py --no-e2e --no-ops --steps 3 --cpu-rows 100000
B2_SORT_CFG=10 step bench_cfg10 400 python bench.py --no-e2e --no-ops --steps 3 --cpu-rows 100000
B2_SORT_CFG=11 step bench_cfg11 400 python bench.py --no-e2e --no-ops --steps 3 --cpu-rows 100000
step bench_extra 900 python bench.py --no-e2e --steps 3
step ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$O/bench_launches.csv" \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-ops --cpu-rows 100000
R=134217728
cap() {  # cap <name> <kernel regex> <skip> <count> <env...> -- <op>
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  local t0=$SECONDS
  env "${envs[@]}" timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s "$skip" -c "$cnt" -f -o "$O/$name" \
    python scripts/profile_ops.py --op "$1" --rows $R > "$O/$name.log" 2>&1
  local rc=$?
  if [ -f "$O/$name.ncu-rep" ]; then
    ncu -i "$O/$name.ncu-rep" --page raw --csv > "$O/${name}_raw.csv" 2>/dev/null
    ncu -i "$O/$name.ncu-rep" --page source --csv > "$O/${name}_src.csv" 2>/dev/null
    rm -f "$O/$name.ncu-rep"
  fi
  echo "ncu_$name exit=$rc secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"
}
# new default: histogram, 4 carry passes, fix-up (warm-up call = 1 + 8 launched onesweep (4 exit at once) + 1)
cap sort_default "onesweep|segment_fix|histogram" 10 10 X=1 -- sort_by_key
cap sort_r1path "onesweep|gather_kernel" 9 9 B2_SORT_HYBRID=0 B2_SORT_CARRY=0 -- sort_by_key
cap join "build_kernel|count_kernel|retrieve_kernel" 3 3 X=1 -- inner_join
cap groupby "groupby_kernel" 1 1 X=1 -- groupby
cap scan "scan_kernel" 1 1 X=1 -- scan
B2_JOIN_RADIX_ROWS=1 step ncu_rjoin_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$O/rjoin_launches.csv" \
  python scripts/profile_ops.py --op inner_join --rows $R
du -sh "$O"
cat "$O/summary.txt"
