"""Summarise an .ncu-rep (ncu --set full) into the metric list the profiles/*_summary.txt files quote.
usage: python tools/ncu_summary.py <file.ncu-rep> [kernel-name substring]"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(head)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if want and want not in name:
            continue
        print("== %s  (grid %s, block %s) ==" % (name[:110], r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]))
        for m in METRICS:
            if m in col:
                print("  %-84s %s %s" % (m, r[col[m]], units[col[m]]))
        st = [(float(r[i].replace(",", "") or 0), h) for h, i in col.items()
              if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
        for v, h in sorted(st, reverse=True)[:9]:
            print("  %-84s %.3f" % (h, v))
        print()


if __name__ == "__main__":
    main()
