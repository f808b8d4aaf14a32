#!/bin/bash
echo "== cost volume tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "cost_volume or stage_seam or cascade" 2>&1 | grep -E "^E  |passed|failed" | head -20
python - <<'PY'
import json
r=json.load(open('gpurun_out/parity_report.json'))
for k,v in sorted(r.items()):
    if 'adaptive_used_pipeline' in v: print(k, v['adaptive_used_pipeline'], v['adaptive_window_miss_permille'])
PY
for wl in dtu tt; do
  for env in "MVSF_WARP_TILE=1" "MVSF_WARP_TILE=2"; do
    echo "== $wl $env"
    env $env timeout 300 python tools/profile_forward.py --workload $wl --iters 2 --breakdown 2>&1 | grep -E "per call|total ms|finest"
  done
done
