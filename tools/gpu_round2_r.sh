#!/bin/bash
# T&T: pipeline kernel vs L1-gather kernel for pass A, and the share of the global fallback
for env in "MVSF_WARP_TILE=1" "MVSF_WARP_TILE=0" "MVSF_WARP_TILE=1 MVSF_WT_DEBUG=4"; do
  echo "== $env"
  env $env timeout 300 python tools/profile_forward.py --workload tt --iters 2 --breakdown 2>&1 | grep -E "per call|total ms"
done
