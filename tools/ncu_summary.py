"""Summarise .ncu-rep files (ncu --set full) into one markdown table per kernel: duration, DRAM bytes,
DRAM %, issue %, occupancy, registers, top stall reasons.   python tools/ncu_summary.py a.ncu-rep b.ncu-rep ..."""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time_us", lambda v: f"{float(v):.1f}"),
    ("dram__bytes_read.sum", "dram_rd_MB", None),
    ("dram__bytes_write.sum", "dram_wr_MB", None),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", lambda v: f"{float(v):.1f}"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%", lambda v: f"{float(v):.1f}"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%", lambda v: f"{float(v):.1f}"),
    ("launch__registers_per_thread", "regs", lambda v: f"{float(v):.0f}"),
    ("launch__grid_size", "grid", lambda v: f"{float(v):.0f}"),
    ("launch__block_size", "block", lambda v: f"{float(v):.0f}"),
    ("smsp__inst_executed.sum", "warp_inst_M", lambda v: f"{float(v) / 1e6:.2f}"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb", lambda v: f"{float(v):.2f}"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb", lambda v: f"{float(v):.2f}"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier", lambda v: f"{float(v):.2f}"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio", lambda v: f"{float(v):.2f}"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg", lambda v: f"{float(v):.2f}"),
]


def to_mb(value, unit):
    v = float(value)
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
    return f"{v * scale:.2f}"


def main():
    print("| kernel | " + " | ".join(n for _, n, _ in WANT) + " |")
    print("|---|" + "---|" * len(WANT))
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        ki = hdr.index("Kernel Name")
        seen = {}
        for r in rows[2:]:
            name = r[ki].split("(")[0].replace("void ", "")
            seen.setdefault(name, r)   # first launch of each kernel in the file
        for name, r in seen.items():
            cells = []
            for metric, _, fmt in WANT:
                if metric not in hdr:
                    cells.append("-")
                    continue
                i = hdr.index(metric)
                cells.append(to_mb(r[i], units[i]) if fmt is None else fmt(r[i]))
            print(f"| {name} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
