#!/bin/bash
# attention A-B: variants of the library under gpurun_variants/ (tools/build_variant.sh)
for K in "$@"; do
  echo "== $K"
  MVSF_LIB_PATH=$PWD/gpurun_variants/libmvsf_b200_$K.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "attention_tensor_core" 2>&1 | tail -1
  python - <<PY
import json
r=json.load(open('gpurun_out/parity_report.json'))
for k,v in sorted(r.items()):
    if k.startswith('attention_N') and k.endswith('plo0') and ('27648' in k or '32640' in k or 'N4000' in k): print(k, {a:float('%.4g'%b) for a,b in v.items()})
PY
done
