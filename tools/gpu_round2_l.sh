#!/bin/bash
for dbg in 0 8 4 12 1 9 2 10 15; do
  echo "== MVSF_WT_DEBUG=$dbg (1: no staging, 2: no window gather, 4: no global fallback, 8: no corr store)"
  MVSF_WT_DEBUG=$dbg timeout 300 python tools/profile_forward.py --iters 2 --breakdown 2>&1 | grep "per call" | python -c "
import sys,json
l=sys.stdin.read().split('per call, last forward: ')[1]
c=json.loads(l)
print([(n.replace('mvsf_warp_corr_',''),t) for n,t in c if 'entropy' in n][2:])"
done
