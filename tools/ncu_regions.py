"""Per-region summary of one kernel of an .ncu-rep (source page): contiguous SASS runs with similar execution counts, with their
share of executed instructions and of stall samples, plus the hottest stall reasons.  Profiling aid, not product code.
usage: python tools/ncu_regions.py report.ncu-rep kernel_id n_envs [--dump out.txt]"""
import csv, io, subprocess, sys
rep, kid, n_env = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = [k for k, l in enumerate(lines) if l.startswith('"Address"')][0]
end = len(lines)
for k in range(start + 1, len(lines)):
    if lines[k].startswith('"Kernel Name"') or lines[k].startswith('"Address"') or lines[k].startswith('"#"') or not lines[k].strip():
        end = k; break
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
tot = sum(float(r["Instructions Executed"] or 0) for r in rows)
tots = sum(float(r["# Samples"] or 0) for r in rows)
print(f"# {lines[0][:120]}\n# instr/env {tot / n_env:.0f}  samples {tots:.0f}")
data = []
for k, r in enumerate(rows):
    ex = float(r["Instructions Executed"] or 0); th = float(r["Thread Instructions Executed"] or 0)
    data.append((k, ex / n_env, th / ex if ex else 0, float(r["# Samples"] or 0) / tots * 100, r["Source"].strip()))
if "--dump" in sys.argv:
    with open(sys.argv[sys.argv.index("--dump") + 1], "w") as f:
        for d in data:
            f.write(f"{d[0]:5d} {d[1]:8.2f} {d[2]:5.1f} {d[3]:6.2f}%  {d[4]}\n")
regions = []; cur = [data[0]]
for d in data[1:]:
    a, b = cur[-1][1], d[1]
    if (a == 0 and b == 0) or (a > 0 and b > 0 and abs(a - b) / max(a, b) < 0.12): cur.append(d)
    else: regions.append(cur); cur = [d]
regions.append(cur)
T = sum(d[1] for d in data)
for r in regions:
    s = sum(d[1] for d in r); smp = sum(d[3] for d in r)
    if s / T > 0.004 or smp > 0.5:
        top = max(r, key=lambda d: d[3])
        print(f"[{r[0][0]:5d}-{r[-1][0]:5d}] len={len(r):4d} x{r[0][1]:7.1f}/env instr={s:7.0f} ({s / T * 100:4.1f}%) samples={smp:5.1f}% lanes={sum(d[2] * d[1] for d in r) / max(s, 1e-9):4.1f}  hot: {top[3]:.1f}% {top[4][:60]}")
