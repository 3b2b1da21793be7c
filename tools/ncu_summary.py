"""Summarise an ncu --set full report as the CSV table kept under profiles/ (one row per profiled launch) and print the DRAM traffic.
usage: ncu_summary.py <report.ncu-rep> <out.csv>"""
import csv
import subprocess
import sys

rep, out = sys.argv[1:3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
idx = [hdr.index(w) for w in want if w in hdr]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        r = list(r)
        r[idx[0]] = r[idx[0]].split("(")[0].replace("void ", "").replace("eslam::", "").replace("<unnamed>::", "")
        w.writerow([r[i] for i in idx])
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "")

    def num(k):
        return float(d[k].replace(",", ""))
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
    print(f"{name[:60]:60s} grid {d['launch__grid_size']:>7s} {num('gpu__time_duration.sum'):9.3f} {units[hdr.index('gpu__time_duration.sum')]}  "
          f"dram {num('dram__bytes_read.sum') * mult[ur] / 1e6:10.1f} + {num('dram__bytes_write.sum') * mult[uw] / 1e6:10.1f} MB  ipc {d.get('sm__inst_executed.avg.per_cycle_elapsed')}")
