#!/bin/bash
mkdir -p gpurun_out
echo "== unit: cost volume"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "cost_volume or stage_seam or cascade" 2>&1 | tail -8
echo "== breakdown"
timeout 300 python tools/profile_forward.py --iters 2 --breakdown > gpurun_out/r2d_breakdown.txt 2>&1; tail -2 gpurun_out/r2d_breakdown.txt
echo "== dsweep"
timeout 600 python bench.py --workload dsweep --steps 6 > gpurun_out/r2d_dsweep.json 2> gpurun_out/r2d_dsweep.err
python -c "
import json; b=json.load(open('gpurun_out/r2d_dsweep.json')); print([(r['D'],r['pass_a_ms'],r['pass_b_ms'],round(r['frac'],4)) for r in b['sweep']])"
echo "== ncu warp_tile"
bash tools/ncu_capture.sh r2d_ncu_warp_tile "warp_tile|warp_corr|corr_aggregate" 12 -- python tools/profile_forward.py --iters 1
python tools/ncu_table.py gpurun_out/r2d_ncu_warp_tile.csv
echo "== fullsize"
timeout 1800 python -m pytest tests/test_gpu_fullsize.py -q 2>&1 | tail -12
cp gpurun_out/parity_report.json gpurun_out/r2d_parity.json
python - <<'PY'
import json
r=json.load(open('gpurun_out/parity_report.json'))
for k,v in sorted(r.items()):
    if k.startswith('fullsize') and ('s3' in k or 's4' in k or 'cascade' in k):
        print(k, {a:(float('%.3g'%b) if isinstance(b,float) else b) for a,b in v.items() if 'prob' in a or 'depth_rel' in a})
PY
du -sh gpurun_out
