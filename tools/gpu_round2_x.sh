#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -q -k "cost_volume or stage_seam or cascade or spill or boundary or install" 2>&1 | tail -2
for wl in dtu tt; do
  echo "== $wl"; timeout 300 python tools/profile_forward.py --workload $wl --iters 2 --breakdown 2>&1 | grep -E "per call|total ms|finest"
done
bash tools/ncu_capture.sh r2_ncu_costvolume "warp_|corr_aggregate|vis_cnn" 16 -- python tools/profile_forward.py --iters 1
python tools/ncu_table.py gpurun_out/r2_ncu_costvolume.csv
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_bench.json 2>/dev/null; python -c "
import json; b=json.load(open('gpurun_out/r2x_bench.json')); print(b['value'], b['e2e']['value'], b['roofline_hbm']['frac'], b['roofline_hbm']['kernel_ms_per_depth_map'], b['gpu_launches_per_step'])"
