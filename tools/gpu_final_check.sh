#!/bin/bash
# end-of-round check on the B200 box: smoke(), the full GPU suite, the default bench line and the reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -3
cp gpurun_out/parity_report.json gpurun_out/r2_final_parity.json
timeout 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -1 gpurun_out/r2_final_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2_final_bench.json'))
print({k:b[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, b['e2e']['value'], b['roofline']['frac'], b['roofline']['launch_ms'], b['roofline_hbm']['frac'], b['roofline_hbm']['kernel_ms_per_depth_map'], b['cpu_baseline']['value'], b['clocks'])
PY
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2>/dev/null; python -c "
import json; b=json.load(open('gpurun_out/r2_final_bench_ref.json')); print('reference arm', b['value'], b['cpu_baseline']['kind'])"
