#!/bin/bash
# first GPU visit of round 2: full GPU test suite (incl. the new full-size parity), bench baseline, launch list, ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2a_pytest.log
cp gpurun_out/parity_report.json gpurun_out/r2a_parity_report.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches.csv python tools/profile_forward.py --iters 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"warp_corr|corr_aggregate|vis_cnn|fmt_smooth|conv3d|linear_tc" -o gpurun_out/r2a_full python tools/profile_forward.py --iters 1 > gpurun_out/r2a_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
tail -5 gpurun_out/r2a_pytest.log
