#!/bin/bash
# attention A-B with the instrumented builds: phase trace of one CTA + time / error of the 27 648-token case
for K in "$@"; do
  echo "== $K"
  MVSF_LIB_PATH=$PWD/gpurun_variants/libmvsf_b200_$K.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -s -k "attention_tensor_core and 27648-0" 2>&1 | grep -E "MMA warp|softmax warp 2 |passed|failed" | head -3
  python - <<PY
import json
r=json.load(open('gpurun_out/parity_report.json'))
v=r['attention_N27648_plo0']; print({a:float('%.4g'%b) for a,b in v.items()})
PY
done
