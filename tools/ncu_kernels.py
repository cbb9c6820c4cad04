"""One line per kernel of an .ncu-rep: duration, instructions, issue utilisation, occupancy, DRAM bytes, top stall reasons.
usage: python tools/ncu_kernels.py report.ncu-rep   (profiling aid)"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
def g(r, name):
    return r[h.index(name)] if name in h else "?"
for r in rows[2:]:
    st = {}
    for k, name in enumerate(h):
        if "pcsamp_warps_issue_stalled" in name and "not_issued" not in name:
            try: st[name.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[k])
            except ValueError: pass
    tot = sum(st.values()) or 1.0
    print(f"{g(r,'ID')} {g(r,'Kernel Name')[:48]} grid={g(r,'launch__grid_size')} t={g(r,'gpu__time_duration.sum')} inst={float(g(r,'smsp__inst_executed.sum'))/float(g(r,'launch__grid_size')):.0f}/cta "
          f"issue={float(g(r,'smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f}% warps={float(g(r,'sm__warps_active.avg.pct_of_peak_sustained_active')):.1f}% "
          f"dram r/w={g(r,'dram__bytes_read.sum')}/{g(r,'dram__bytes_write.sum')} {g(r,'dram__bytes_read.sum.unit') if 'dram__bytes_read.sum.unit' in h else ''} L1hit={float(g(r,'l1tex__t_sector_hit_rate.pct')):.0f}% L2hit={float(g(r,'lts__t_sector_hit_rate.pct')):.0f}% regs={g(r,'launch__registers_per_thread')}")
    print("     stalls: " + ", ".join(f"{k} {v / tot * 100:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
