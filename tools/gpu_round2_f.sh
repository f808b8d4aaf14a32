#!/bin/bash
mkdir -p gpurun_out
echo "== full GPU test suite"
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12
cp gpurun_out/parity_report.json gpurun_out/r2f_parity.json
python - <<'PY'
import json
r=json.load(open('gpurun_out/parity_report.json'))
for k,v in sorted(r.items()):
    if k.startswith('attention_N27648') or k.startswith('attention_N32640') or 'install' in k or 'batch2' in k or 'pe_on' in k or 'two_gather' in k or 'python' in k: print(k, {a:(round(b,8) if isinstance(b,float) else b) for a,b in v.items()})
PY
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -2 gpurun_out/r2f_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2f_bench.json'))
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e'], {k:b['roofline'][k] for k in ('frac','launch_ms')}, {k:b['roofline_hbm'][k] for k in ('frac','kernel_ms_per_depth_map')}); print(b['kernel_ms_per_depth_map']); print(b['cpu_baseline'])
PY
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; python -c "
import json; b=json.load(open('gpurun_out/r2f_bench_ref.json')); print(b['value'], b['cpu_baseline'])"
du -sh gpurun_out
