#!/usr/bin/env bash
# Round 2, final GPU call (1 GPU): full GPU suite as the driver runs it, default bench line, ncu launch list of the bench command and
# `--set full` captures of the shipped kernels at 2^27 rows (summaries by scripts/ncu_summary.py).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2c11
mkdir -p "$O"
step() { local name=$1 to=$2; shift 2; local t0=$SECONDS; timeout "$to" "$@" > "$O/$name.log" 2>&1; echo "$name exit=$? secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"; }
: > "$O/summary.txt"
step tests_all 900 python -m pytest tests -q -m gpu -x --durations=10
step bench 500 python bench.py --steps 10 --warmup 3
step ncu_launches 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$O/bench_launches.csv" \
  python bench.py --steps 2 --warmup 3 --no-e2e --cpu-rows 100000
R=134217728
cap() {  # cap <name> <kernel regex> <skip> <count> <row of the summary> <op>
  local name=$1 rx=$2 skip=$3 cnt=$4 row=$5 op=$6
  local t0=$SECONDS
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s "$skip" -c "$cnt" -f -o "$O/$name" \
    python scripts/profile_ops.py --op "$op" --rows $R > "$O/$name.log" 2>&1
  local rc=$?
  if [ -f "$O/$name.ncu-rep" ]; then
    ncu -i "$O/$name.ncu-rep" --page raw --csv > "$O/${name}_raw.csv" 2>/dev/null
    ncu -i "$O/$name.ncu-rep" --page source --csv > "$O/${name}_src.csv" 2>/dev/null
    python scripts/ncu_summary.py "$O/${name}_raw.csv" "$O/${name}_src.csv" "$row" > "$O/${name}_summary.txt" 2>&1
    rm -f "$O/$name.ncu-rep" "$O/${name}_src.csv"
  fi
  echo "ncu_$name exit=$rc secs=$((SECONDS - t0))" | tee -a "$O/summary.txt"
}
# rows of the raw CSV: 0 header, 1 units, 2.. launches
cap onesweep_carry "onesweep_kernel" 5 1 2 sort_by_key
cap segment_fix "segment_fix_kernel" 2 2 2 sort_by_key
cap rj_join "rj_join_kernel" 1 1 2 inner_join
cap pgb_agg "pgb_agg_kernel" 1 1 2 groupby
cap gb_est_pass "onesweep_kernel" 1 1 2 groupby
cap histogram "histogram_kernel" 2 1 2 sort_by_key
du -sh "$O"
tail -12 "$O/tests_all.log"
grep -h '"metric"' "$O/bench.log" | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print(d['ms_per_step'], d['e2e'].get('ms_per_step'), d['roofline']['frac'], d['roofline']['whole_op']['frac_contract'])
for k, v in d['ops'].items(): print(k, round(v.get('ms', -1), 2), round(v.get('roofline', {}).get('frac', 0), 3), v.get('phases_ms'))
"
for f in onesweep_carry segment_fix rj_join pgb_agg gb_est_pass histogram; do echo "== $f"; head -11 "$O/${f}_summary.txt"; done
cat "$O/summary.txt"
