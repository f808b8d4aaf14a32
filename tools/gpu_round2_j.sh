#!/bin/bash
mkdir -p gpurun_out
echo "== tests"
timeout 1800 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_parity.py tests/test_gpu_boundary.py -q 2>&1 | tail -5
python - <<'PY'
import json
for l in open('gpurun_out/tcgen05_report.jsonl'):
    d=json.loads(l)
    if d['gelu']: print('linear gelu', d['M'], d['N'], d['K'], 'err', d['tc_vs_f64'], 'torch f32 err', d['torch_f32_vs_f64'])
r=json.load(open('gpurun_out/parity_report.json'))
for k,v in sorted(r.items()):
    if k.startswith('fmt_') or k.startswith('cascade_hotpath_v3_96x128_final'): print(k, {a:(float('%.3g'%b) if isinstance(b,float) else b) for a,b in v.items()})
PY
echo "== breakdown"
timeout 300 python tools/profile_forward.py --iters 2 --breakdown > gpurun_out/r2j_breakdown.txt 2>&1; grep -E "costreg_tr|fmt_forward" -A1 gpurun_out/r2j_breakdown.txt | grep -E "mvsf|ms_per"; grep "total ms" gpurun_out/r2j_breakdown.txt
echo "== bench dtu"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2j_bench.json'))
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], {k:b['roofline'][k] for k in ('frac','launch_ms')}, {k:b['roofline_hbm'][k] for k in ('frac','traffic')}); print(b['kernel_ms_per_depth_map'])
PY
echo "== bench tt"
timeout 600 python bench.py --workload tt --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_tt.json 2> gpurun_out/r2j_bench_tt.err; tail -2 gpurun_out/r2j_bench_tt.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2j_bench_tt.json'))
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value']); print(b['kernel_ms_per_depth_map'])
PY
