"""Summarise an `ncu --page raw --csv` export: one line per kernel launch with the metrics that decide whether a kernel is
HBM-, issue- or latency-bound.  python tools/ncu_summary.py raw.csv [max_per_kernel]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 3
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
COLS = [("t_us", "gpu__time_duration.sum", 1.0), ("rdMB", "dram__bytes_read.sum", 1.0), ("wrMB", "dram__bytes_write.sum", 1.0),
        ("dram%", "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", 1.0), ("lts%", "LTS.TriageCompute.lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("issue%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1.0), ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0), ("regs", "launch__registers_per_thread", 1.0),
        ("st_long", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0),
        ("st_short", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 1.0),
        ("st_bar", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1.0),
        ("st_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", 1.0),
        ("st_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", 1.0),
        ("st_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", 1.0),
        ("bankc", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1.0)]
print("kernel".ljust(26), "grid".ljust(14), " ".join(n.rjust(8) for n, _, _ in COLS))
seen = collections.Counter()
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    key = (name, r[idx["Grid Size"]])
    seen[key] += 1
    if seen[key] > cap:
        continue
    vals = []
    for n, k, _ in COLS:
        try:
            v = float(r[idx[k]].replace(",", ""))
            if n == "t_us" and units[idx[k]] == "ns":
                v /= 1e3
            if n in ("rdMB", "wrMB"):
                u = units[idx[k]]
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            vals.append(f"{v:8.1f}")
        except Exception:
            vals.append("       -")
    print(name[:26].ljust(26), r[idx["Grid Size"]].ljust(14), " ".join(vals))
