#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 4 1 5 2 6 7; do
  echo "== MVSF_WT_DEBUG=$dbg (1: no staging, 2: no window gather, 4: no global fallback)"
  MVSF_WT_DEBUG=$dbg timeout 300 python tools/profile_forward.py --iters 2 --breakdown 2>&1 | grep "per call" | python -c "
import sys,json
l=sys.stdin.read().split('per call, last forward: ')[1]
c=json.loads(l)
print([(n.replace('mvsf_warp_corr_',''),t) for n,t in c if 'entropy' in n or 'aggregate' in n][4:])"
done
echo "== attention poly test"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "attention" 2>&1 | tail -3
python - <<'PY'
import json
r=json.load(open('gpurun_out/parity_report.json'))
for k,v in sorted(r.items()):
    if k.startswith('attention_N27648') or k.startswith('attention_N32640'): print(k, {a:(round(b,8) if isinstance(b,float) else b) for a,b in v.items()})
PY
timeout 300 python tools/profile_forward.py --iters 2 --breakdown 2>&1 | grep -A2 costreg_tr
