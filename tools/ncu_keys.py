"""Print the handful of ncu raw-page metrics used in profiles/README.md from a .ncu-rep (run here, no GPU needed)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(d["Kernel Name"][:70], "grid", d["Grid Size"], "block", d["Block Size"])
        for k in hdr:
            short = k.split(".", 2)[-1] if k.count(".") >= 2 and k.split(".")[1].startswith("Triage") else k
            if any(k.endswith(s) or short == s for s in KEYS) or "warp_issue_stalled" in k and k.endswith("_per_warp_active.pct"):
                try:
                    if float(d[k].replace(",", "")) < 0.5 and "stalled" in k:
                        continue
                except ValueError:
                    pass
                print(f"   {k} = {d[k]} {u[k]}")


if __name__ == "__main__":
    main()
