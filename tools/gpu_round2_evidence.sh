#!/bin/bash
# ncu evidence of the kernels as shipped: CSV exports only (the .ncu-rep files stay on the box)
mkdir -p gpurun_out
bash tools/ncu_capture.sh r2_ncu_costvolume "warp_|corr_aggregate|vis_cnn" 16 -- python tools/profile_forward.py --iters 1
bash tools/ncu_capture.sh r2_ncu_fmt "fmt_smooth|reduce1x1|linattn|kv_partial|layernorm64" 12 -- python tools/profile_forward.py --iters 1
bash tools/ncu_capture.sh r2_ncu_linear "linear_tc" 10 -s 20 -- python tools/profile_forward.py --iters 1
bash tools/ncu_capture.sh r2_ncu_conv3d "conv3d" 27 -- python tools/profile_forward.py --iters 1
# attention: --set full fails to launch under ncu's patching (r1 finding); the light sections work
timeout 600 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats --clock-control none -k regex:attention_fa -c 2 -f -o /tmp/r2_ncu_attention python tools/profile_forward.py --iters 1 > gpurun_out/r2_ncu_attention.log 2>&1
ncu -i /tmp/r2_ncu_attention.ncu-rep --page raw --csv > gpurun_out/r2_ncu_attention.csv 2>> gpurun_out/r2_ncu_attention.log
for f in costvolume fmt linear conv3d attention; do python tools/ncu_table.py gpurun_out/r2_ncu_$f.csv; done
du -sh gpurun_out
