#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "prefetching or cost_volume" 2>&1 | tail -3
for st in 1 2; do
timeout 600 python bench.py --steps 12 --warmup 3 --streams $st --no-cpu-baseline > gpurun_out/r2t_bench_s$st.json 2> gpurun_out/r2t_bench_s$st.err; tail -2 gpurun_out/r2t_bench_s$st.err
python - <<PY
import json
b=json.load(open('gpurun_out/r2t_bench_s$st.json'))
print('streams $st', {k:b[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, b['e2e']['value'], b['config']['compute_streams'], b['clocks'])
PY
done
timeout 600 python bench.py --workload tt --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench_tt.json 2>/dev/null; python -c "
import json; b=json.load(open('gpurun_out/r2t_bench_tt.json')); print('tt', b['value'], b['ms_per_step'], b['e2e']['value'])"
