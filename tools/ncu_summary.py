"""Compact summary (metric,unit,value) of one kernel of a .ncu-rep, the format of profiles/*_ncu_full_summary.csv.
Usage: python tools/ncu_summary.py report.ncu-rep > profiles/name.csv   (run here, no GPU needed)."""
import csv
import subprocess
import sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    w = csv.writer(sys.stdout)
    w.writerow(["metric", "unit", "value"])
    d = {k: (u, v) for k, u, v in zip(hdr, units, vals)}
    for k in ("Kernel Name", "Block Size", "Grid Size"):
        w.writerow([k, "", d[k][1]])
    for k in sorted(d):
        stall = k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")
        if k in KEEP or stall:
            if d[k][1] != "":
                w.writerow([k, d[k][0], d[k][1]])


if __name__ == "__main__":
    main()
