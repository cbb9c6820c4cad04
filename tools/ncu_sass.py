"""Dump the SASS page of an .ncu-rep as 'idx  exec/env  lanes  samples  instruction' (profiling aid; not product code).
usage: python tools/ncu_sass.py report.ncu-rep n_envs [kernel-substring] > out.txt"""
import csv, io, subprocess, sys
rep, n_env = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(k for k, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(float(r["Instructions Executed"] or 0) for r in rows)
tots = sum(float(r["# Samples"] or 0) for r in rows)
print(f"# instr/env {tot / n_env:.0f}  samples {tots:.0f}")
for k, r in enumerate(rows):
    ex = float(r["Instructions Executed"] or 0)
    th = float(r["Thread Instructions Executed"] or 0)
    print(f"{k:5d} {ex / n_env:8.2f} {th / ex if ex else 0:5.1f} {float(r['# Samples'] or 0) / tots * 100:6.2f}%  {r['Source'].strip()}")
