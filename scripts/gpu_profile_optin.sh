#!/usr/bin/env bash
# ncu evidence for the opt-in paths once they are validated (second or third GPU call of a round):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_profile_optin.sh'
# Launch lists (gpu__time_duration.sum) of every operation, then --set full captures of the kernels that matter.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/optin
mkdir -p "$O"
ROWS=${ROWS:-134217728}
list() {  # list <name> <env...> -- <op>
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$O/${name}_launches.csv" \
    python scripts/profile_ops.py --op "$1" --rows "$ROWS" > "$O/${name}_launches.log" 2>&1
  echo "$name launches exit=$?" | tee -a "$O/summary.txt"
}
full() {  # full <name> <kernel regex> <skip> <env...> -- <op>
  local name=$1 rx=$2 skip=$3; shift 3
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s "$skip" -c 2 -f -o "$O/${name}_full" \
    python scripts/profile_ops.py --op "$1" --rows "$ROWS" > "$O/${name}_full.log" 2>&1
  echo "$name full exit=$?" | tee -a "$O/summary.txt"
}
: > "$O/summary.txt"
list default_sort X=1 -- sort_by_key
list alias_sort B2_SORT_ALIAS=1 -- sort_by_key
list carry_sort B2_SORT_CARRY=1 -- sort_by_key_payload
list hash_join X=1 -- inner_join
list radix_join B2_JOIN_RADIX_ROWS=1 -- inner_join
list groupby X=1 -- groupby
list scan X=1 -- scan
full alias_onesweep onesweep 8 B2_SORT_ALIAS=1 -- sort_by_key
full carry_onesweep onesweep 8 B2_SORT_CARRY=1 -- sort_by_key_payload
full radix_join_kernel rj_join 2 B2_JOIN_RADIX_ROWS=1 -- inner_join
full groupby_kernel groupby_kernel 1 X=1 -- groupby
cat "$O/summary.txt"
