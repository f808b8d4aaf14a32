#!/bin/bash
echo "== vis unit tests, XLO=1 (default)"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "vis_cnn or cost_volume" 2>&1 | tail -2
echo "== vis unit tests, XLO=0 (errors only)"
MVSF_VIS_XLO=0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "vis_cnn or cost_volume" 2>&1 | tail -2
python - <<'PY'
import json
r=json.load(open('gpurun_out/parity_report.json'))
print('vis errors XLO=0:', {k:(v.get('abs') or v.get('vis_isolated')) for k,v in sorted(r.items()) if k.startswith('vis_cnn') or k.startswith('cost_volume')})
PY
echo "== fullsize dtu-white teacher-forced, XLO=0"
MVSF_VIS_XLO=0 timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -k "dtu-white" 2>&1 | tail -3
python - <<'PY'
import json
r=json.load(open('gpurun_out/parity_report.json'))
for k,v in sorted(r.items()):
    if k.startswith('fullsize_dtu_white_teacher'): print(k, {a:float('%.3g'%b) for a,b in v.items() if a in ('vs_hom64_prob','vs_hom64_vis','vs_hom64_volume','vs_hom64_logits','floor_prob','vs_plain_prob')})
PY
for x in 1 0; do
echo "== breakdown XLO=$x"
MVSF_VIS_XLO=$x timeout 300 python tools/profile_forward.py --iters 2 --breakdown 2>&1 | grep -E "vis_cnn\"" -A1 | grep ms_per
done
