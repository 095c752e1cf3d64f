#!/usr/bin/env bash
# First GPU call of a round: everything that was written without hardware gets its verdict in ONE gpurun call.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_first_call.sh'
# Every step has its own timeout and log under gpurun_out/first/; a failing step does not stop the rest.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/first
mkdir -p "$O"
step() {  # step <name> <timeout-seconds> <command...>
  local name=$1 to=$2; shift 2
  local t0=$SECONDS
  timeout "$to" "$@" > "$O/$name.log" 2>&1
  echo "$name exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"
}
: > "$O/summary.txt"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > "$O/gpu.txt" 2>&1

# 1. the validated suite first, then the files that never ran on hardware (each on its own so that one cannot hide another)
NEW="tests/test_zzzz_match_context.py tests/test_zzzz_var_std.py tests/test_zzzz_segmented_sort.py tests/test_zzzz_rank.py tests/test_zzzz_partitioning.py"
IGN=""; for f in $NEW; do IGN="$IGN --ignore $f"; done
step tests_validated 1500 python -m pytest tests -q -m gpu -x --ignore tests/test_zz_experimental_gpu.py \
  --ignore tests/test_zz_full_size_gpu.py --ignore tests/test_zzz_cpp_api.py --ignore tests/test_zy_more_golden.py $IGN
# written on the emulator, never on hardware: match context / partitioned probes, VAR / STD, segmented sort / top-k, rank, partitioning
step tests_emulator_born 900 python -m pytest $NEW -q -m gpu
step tests_more_golden 300 python -m pytest tests/test_zy_more_golden.py -q -m gpu
step tests_cpp_api 300 python -m pytest tests/test_zzz_cpp_api.py -q -m gpu
B2_RUN_EXPERIMENTAL=1 step tests_experimental 1000 python -m pytest tests/test_zz_experimental_gpu.py -q -m gpu -rxX
step tests_full_size 900 python -m pytest tests/test_zz_full_size_gpu.py -q -m gpu

# 2. headline bench, the reference arm, and the secondary ops (keys-only / aliased sort, join, groupby, scan, reduce)
step bench 600 python bench.py
step bench_reference 600 python bench.py --impl reference --steps 3 --warmup 1
step bench_extra 900 python bench.py --extra --no-e2e
B2_SORT_ALIAS=1 step bench_alias 600 python bench.py --no-e2e
B2_SORT_CARRY=1 step bench_carry 600 python bench.py --no-e2e

# 3. ncu: launch list of the bench command itself (shares of the step), then one full capture of the pass kernel
step ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$O/bench_launches.csv" \
  python bench.py --steps 2 --warmup 1 --no-e2e
step ncu_onesweep 900 ncu --set full --clock-control none --import-source on -k regex:onesweep -s 8 -c 2 -f -o "$O/onesweep_full" \
  python scripts/profile_sort.py --rows 134217728
cat "$O/summary.txt"
