"""Key numbers + stall breakdown of one .ncu-rep (profiling aid). usage: python tools/ncu_summary.py rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines())); h, v = r[0], r[2]
g = lambda k: v[h.index(k)] if k in h else None
for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]:
    print(f"{k:70s} {g(k)}")
items = [(float(x), k[len('smsp__pcsamp_warps_issue_stalled_'):]) for k, x in zip(h, v)
         if k.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in k and x not in ("", None)]
tot = sum(x for x, _ in items)
print("stalls: " + ", ".join(f"{k} {x / tot * 100:.1f}%" for x, k in sorted(items, reverse=True)[:9]))
