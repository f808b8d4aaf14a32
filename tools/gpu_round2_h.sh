#!/bin/bash
mkdir -p gpurun_out
echo "== tests: linear / fmt / transformer / attention"
timeout 1200 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_parity.py -q -k "linear or fmt or costreg or attention" 2>&1 | tail -4
python - <<'PY'
import json
r=json.load(open('gpurun_out/parity_report.json'))
for k,v in sorted(r.items()):
    if k.startswith('attention_N27648'): print(k, {a:(round(b,8) if isinstance(b,float) else b) for a,b in v.items()})
PY
for plo in 0 2; do
echo "== breakdown MVSF_ATTENTION_PLO=$plo"
MVSF_ATTENTION_PLO=$plo timeout 300 python tools/profile_forward.py --iters 2 --breakdown > gpurun_out/r2h_breakdown_$plo.txt 2>&1; grep -E "costreg_tr|fmt_forward" -A1 gpurun_out/r2h_breakdown_$plo.txt | grep -E "mvsf|ms_per"; grep "total ms" gpurun_out/r2h_breakdown_$plo.txt
done
echo "== fullsize dtu stage 1 with one-product scores"
MVSF_ATTENTION_PLO=2 timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -k "dtu-image and (cascade or forced-1 or 1])" 2>&1 | tail -4
python - <<'PY'
import json
r=json.load(open('gpurun_out/parity_report.json'))
for k,v in sorted(r.items()):
    if k.startswith('fullsize') and ('s1' in k or 'cascade' in k): print(k, {a:(float('%.3g'%b) if isinstance(b,float) else b) for a,b in v.items() if 'prob' in a or 'depth_rel' in a or 'logits' in a})
PY
