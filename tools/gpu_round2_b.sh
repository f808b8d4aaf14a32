#!/bin/bash
# keep gpurun_out small (<64 MiB): summaries only
mkdir -p gpurun_out
echo "== fullsize, L1 organisation (round-1 kernels)"
MVSF_WARP_TILE=0 timeout 900 python -m pytest tests/test_gpu_fullsize.py -q 2>&1 | tail -15
cp gpurun_out/parity_report.json gpurun_out/r2b_parity_fullsize_l1.json
echo "== cost volume unit tests (window kernels)"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "cost_volume or stage_seam or cascade" 2>&1 | tail -15
cp gpurun_out/parity_report.json gpurun_out/r2b_parity_unit.json
echo "== fullsize, window kernels"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q 2>&1 | tail -15
cp gpurun_out/parity_report.json gpurun_out/r2b_parity_fullsize_tile.json
echo "== breakdown"
MVSF_WARP_TILE=0 timeout 300 python tools/profile_forward.py --iters 2 --breakdown > gpurun_out/r2b_breakdown_l1.txt 2>&1
timeout 300 python tools/profile_forward.py --iters 2 --breakdown > gpurun_out/r2b_breakdown_tile.txt 2>&1
tail -3 gpurun_out/r2b_breakdown_l1.txt; tail -3 gpurun_out/r2b_breakdown_tile.txt
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
cat gpurun_out/r2b_bench.json | head -c 3000
timeout 600 python bench.py --workload dsweep --steps 6 > gpurun_out/r2b_dsweep.json 2> gpurun_out/r2b_dsweep.err
cat gpurun_out/r2b_dsweep.json | head -c 2500; tail -3 gpurun_out/r2b_dsweep.err
du -sh gpurun_out
