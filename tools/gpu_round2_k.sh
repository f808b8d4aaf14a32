#!/bin/bash
mkdir -p gpurun_out
echo "== unit: cost volume, seams, cascade"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -q -k "cost_volume or stage_seam or cascade or two_gather or install" 2>&1 | tail -8
echo "== breakdown"
timeout 300 python tools/profile_forward.py --iters 2 --breakdown > gpurun_out/r2k_breakdown.txt 2>&1; tail -2 gpurun_out/r2k_breakdown.txt
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -3 gpurun_out/r2k_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2k_bench.json'))
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], {k:b['roofline_hbm'][k] for k in ('frac','kernel_ms_per_depth_map')}); print(b['kernel_ms_per_depth_map'])
PY
