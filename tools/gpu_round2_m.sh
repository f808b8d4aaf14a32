#!/bin/bash
echo "== unit"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "cost_volume" 2>&1 | tail -3
for cfg in "3 2" "2 3"; do
  set -- $cfg
  echo "== BLOCKS8=$1 NBUF8=$2"
  MVSF_EXTRA_NVCC_FLAGS="-DMVSF_PS_BLOCKS8=$1 -DMVSF_PS_NBUF8=$2" python -m mvsformerplusplus_b200.build > /dev/null 2>&1
  for dbg in 0 12; do
  MVSF_EXTRA_NVCC_FLAGS="-DMVSF_PS_BLOCKS8=$1 -DMVSF_PS_NBUF8=$2" MVSF_WT_DEBUG=$dbg timeout 300 python tools/profile_forward.py --iters 2 --breakdown 2>&1 | grep "per call" | python -c "
import sys,json
l=sys.stdin.read().split('per call, last forward: ')[1]
c=json.loads(l)
print('dbg', $dbg, [(n.replace('mvsf_warp_corr_',''),t) for n,t in c if 'entropy' in n][2:])"
  done
done
python -m mvsformerplusplus_b200.build > /dev/null 2>&1
