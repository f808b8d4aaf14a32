#!/bin/bash
# round-2 final evidence: full GPU suite, bench (DTU + T&T), launch list, ncu of the shipped kernels (CSV exports only)
mkdir -p gpurun_out
echo "== full GPU test suite"
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6
cp gpurun_out/parity_report.json gpurun_out/r2q_parity.json
echo "== bench DTU"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -2 gpurun_out/r2q_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2q_bench.json'))
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e'], {k:b['roofline'][k] for k in ('frac','launch_ms')}, {k:b['roofline_hbm'][k] for k in ('frac','kernel_ms_per_depth_map')}); print(b['kernel_ms_per_depth_map']); print(b['cpu_baseline'])
PY
echo "== bench T&T"
timeout 600 python bench.py --workload tt --steps 8 --warmup 3 > gpurun_out/r2q_bench_tt.json 2> gpurun_out/r2q_bench_tt.err; python -c "
import json; b=json.load(open('gpurun_out/r2q_bench_tt.json')); print(b['value'], b['ms_per_step'], b['e2e'])"
echo "== breakdown + launch list"
timeout 300 python tools/profile_forward.py --iters 2 --breakdown > gpurun_out/r2q_breakdown.txt 2>&1; tail -3 gpurun_out/r2q_breakdown.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2q_launches.csv python tools/profile_forward.py --iters 1 > /dev/null 2>&1
echo "== ncu --set full of the shipped kernels"
bash tools/gpu_round2_evidence.sh 2>&1 | tail -60
du -sh gpurun_out
