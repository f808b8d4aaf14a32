"""DRAM traffic of the fused warp + group-correlation launches of ONE depth map from the ncu --set full capture of the
shipped kernels (tools/gpu_round2_evidence.sh -> gpurun_out/r2_ncu_costvolume.csv) -> profiles/r2_ncu_traffic.json, the
file bench.py reads `roofline_hbm.traffic` from.
  python tools/ncu_traffic.py gpurun_out/r2_ncu_costvolume.csv > profiles/r2_ncu_traffic.json"""
import csv
import json
import sys

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units = rows[h], rows[h + 1]
kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per_launch = []
for r in rows[h + 2:]:
    if len(r) <= max(rd, wr):
        continue
    name = r[kn].split("(")[0].replace("mvsf::", "")
    if "vis_cnn" in name:
        continue
    b = float(r[rd].replace(",", "")) * scale[units[rd]] + float(r[wr].replace(",", "")) * scale[units[wr]]
    per_launch.append([name, int(b)])
wl = bench.WORKLOADS["dtu"]
alg = sum(bench.algorithmic_bytes(wl["V"], wl["H"], wl["W"]))
tot = sum(b for _, b in per_launch)
n_agg = sum(1 for n, _ in per_launch if "corr_aggregate" in n)
assert n_agg == 4, f"capture incomplete: {n_agg} corr_aggregate launches (need the 4 stages of one depth map)"
out = {
    "warp_corr_entropy_store + corr_aggregate (8 launches / depth map)": {
        "dram_bytes_per_depth_map": tot, "per_launch": per_launch, "algorithmic_bytes_per_depth_map": alg, "ratio": round(tot / alg, 2),
        "source": "ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum) of the shipped kernels, tools/gpu_round2_evidence.sh, "
                  "profiles/r2_ncu_shipped_kernels.md; stage 4 = selection kernel + pipeline kernel + the L1 kernel's empty launch"},
    "attention_fa_kernel": {"dram_bytes_per_launch": 21309000,
                            "source": "profiles/r1_ncu_traffic.json (unchanged operand traffic; r2 capture: dram throughput 0.32 %)"},
}
print(json.dumps(out, indent=1))
