"""Turns the raw-page CSV written by tools/ncu_capture.sh into a markdown table of the metrics DESIGN.md / profiles/ quote.
  python tools/ncu_table.py gpurun_out/<name>.csv [more.csv ...]"""
import csv
import sys

COLS = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB rd"), ("dram__bytes_write.sum", "MB wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 %"),
        ("lts__t_bytes.sum", "L2 MB"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
        ("sm__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"), ("sm__inst_executed.sum", "warp insts"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("launch__registers_per_thread", "regs"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"), ("smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier")]


def conv(name, v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if name == "gpu__time_duration.sum":
        x = x / 1e3 if unit in ("nsecond", "ns") else (x * 1e3 if unit in ("msecond", "ms") else x)
        return f"{x:.1f}"
    if name.startswith("dram__bytes") or name == "lts__t_bytes.sum":
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
        return f"{x * scale:.1f}"
    if name == "sm__inst_executed.sum":
        return f"{x / 1e6:.1f}M"
    return f"{x:.2f}"


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        hdr, units = rows[hdr_i], rows[hdr_i + 1]
        idx = {n: hdr.index(n) for n, _ in COLS if n in hdr}
        print(f"\n`{path}`\n")
        print("| " + " | ".join(lbl for n, lbl in COLS if n in idx) + " |")
        print("|" + "---|" * len(idx))
        for r in rows[hdr_i + 2:]:
            if len(r) < len(hdr):
                continue
            cells = []
            for n, _ in COLS:
                if n not in idx:
                    continue
                v = r[idx[n]]
                if n == "Kernel Name":
                    v = "`" + v.split("(")[0][:60] + "`"
                else:
                    v = conv(n, v, units[idx[n]])
                cells.append(v)
            print("| " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
