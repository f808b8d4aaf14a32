#!/bin/bash
mkdir -p gpurun_out
echo "== unit: cost volume + attention"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "cost_volume or attention" 2>&1 | tail -8
python - <<'PY'
import json
r=json.load(open('gpurun_out/parity_report.json'))
for k,v in sorted(r.items()):
    if k.startswith('attention'): print(k, {a:(round(b,8) if isinstance(b,float) else b) for a,b in v.items()})
PY
echo "== fullsize"
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -q 2>&1 | tail -25
cp gpurun_out/parity_report.json gpurun_out/r2c_parity.json
python - <<'PY'
import json
r=json.load(open('gpurun_out/parity_report.json'))
for k,v in sorted(r.items()):
    if k.startswith('fullsize'): print(k, {a:(float('%.3g'%b) if isinstance(b,float) else b) for a,b in v.items()})
PY
echo "== breakdown (tile path, pass B 3 CTAs/SM)"
timeout 300 python tools/profile_forward.py --iters 2 --breakdown > gpurun_out/r2c_breakdown_b3.txt 2>&1; tail -2 gpurun_out/r2c_breakdown_b3.txt
echo "== breakdown attention PLO=1"
MVSF_ATTENTION_PLO=1 timeout 300 python tools/profile_forward.py --iters 2 --breakdown > gpurun_out/r2c_breakdown_plo1.txt 2>&1; grep -E "costreg_tr|total" gpurun_out/r2c_breakdown_plo1.txt; grep -E "costreg_tr" -A2 gpurun_out/r2c_breakdown_b3.txt
echo "== rebuild pass B 2 CTAs/SM"
MVSF_EXTRA_NVCC_FLAGS="-DMVSF_WT_PASSB_BLOCKS=2" python -m mvsformerplusplus_b200.build > /dev/null 2>&1
MVSF_EXTRA_NVCC_FLAGS="-DMVSF_WT_PASSB_BLOCKS=2" timeout 300 python tools/profile_forward.py --iters 2 --breakdown > gpurun_out/r2c_breakdown_b2.txt 2>&1; tail -2 gpurun_out/r2c_breakdown_b2.txt
python -m mvsformerplusplus_b200.build > /dev/null 2>&1
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2c_bench.json'))
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e'], {k:b['roofline'][k] for k in ('frac','launch_ms')}, {k:b['roofline_hbm'][k] for k in ('frac','kernel_ms_per_depth_map')}); print(b['kernel_ms_per_depth_map'])
PY
timeout 600 python bench.py --workload dsweep --steps 6 > gpurun_out/r2c_dsweep.json 2> gpurun_out/r2c_dsweep.err
python -c "
import json; b=json.load(open('gpurun_out/r2c_dsweep.json')); print([(r['D'],r['pass_a_ms'],r['pass_b_ms'],round(r['frac'],4)) for r in b['sweep']])"
du -sh gpurun_out
