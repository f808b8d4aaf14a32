"""Times the CPU reference arm of bench.py for a few thread counts (run on the GPU box to pick the default)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

wl = bench.WORKLOADS["dtu"]
for t in [int(x) for x in (sys.argv[1:] or ["16", "32", "64"])]:
    dt = bench.cpu_reference_pass(wl, t)
    print(f"threads {t}: {dt:.1f} s per depth map", flush=True)
