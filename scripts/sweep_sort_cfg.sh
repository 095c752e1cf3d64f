#!/usr/bin/env bash
# One-sweep tile-shape sweep (B2_SORT_CFG 0..12, see radix_sort.cu::run_radix) on the headline workload.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/sweep_sort_cfg.sh'
# Prints one line per configuration: ms per sort_by_key, average one-sweep launch, fraction of the HBM peak.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/sweep
mkdir -p "$O"
CFGS="0 1 2 3 4 5 6 7 8 9 10 11"
# cfg 12 (bulk async copy + mbarrier) has never run on hardware: it joins the sweep only after a small sort finished and verified
if B2_SORT_CFG=12 timeout 120 python - > "$O/cfg12_precheck.log" 2>&1 <<'PY'
import numpy as np, torch
import cudf_b200.pylibcudf as plc
k = torch.randint(-2**62, 2**62, (3_000_000,), dtype=torch.int64, device="cuda")
o = plc.sorting.sort_by_key(plc.Table([plc.Column.from_torch(k)]), plc.Table([plc.Column.from_torch(k)]), [0], []).columns()[0].to_torch()
assert bool((o == torch.sort(k).values).all())
print("CFG12_OK")
PY
then CFGS="$CFGS 12"; else echo "cfg 12: precheck failed or timed out, skipped (see $O/cfg12_precheck.log)"; fi
for cfg in $CFGS; do
  B2_SORT_CFG=$cfg timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --cpu-rows 100000 > "$O/cfg$cfg.json" 2> "$O/cfg$cfg.err"
  python - "$O/cfg$cfg.json" "$cfg" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"cfg {sys.argv[2]}: {d['ms_per_step']:.2f} ms/step  onesweep {r['avg_launch_ms']:.3f} ms  frac {r['frac']:.3f}  "
          f"hist {r['other_kernels_ms_per_step']['histogram']:.2f} gather {r['other_kernels_ms_per_step']['gather']:.2f}")
except Exception as ex:
    print(f"cfg {sys.argv[2]}: failed ({ex!r})")
PY
done | tee "$O/summary.txt"
