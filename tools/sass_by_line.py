"""Static SASS instruction count of one kernel per OUTERMOST source line (nvdisasm -gi), bucketed by phase markers.
usage: python tools/sass_by_line.py lib.so mangled_kernel_name [kernels.cuh]   (profiling aid)"""
import bisect, collections, os, re, subprocess, sys, tempfile
so, kern = sys.argv[1], sys.argv[2]
src = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("-") else os.path.join(os.path.dirname(__file__), "..", "marl_llm_b200", "csrc", "swarm_kernels.cuh")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
txt = []
for cub in sorted(f for f in os.listdir(d) if f.endswith(".cubin")):
    t = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
    if ("\n" + kern + ":") in t:
        txt = t.splitlines(); break
on = False; cur = None; pend = None; cnt = collections.Counter(); seq = []
for l in txt:
    if l.startswith(kern + ":"): on = True; continue
    if on and l.startswith("//---------------------"): break
    if not on: continue
    m = re.search(r'//## File "(.*?)", line (\d+)(.*)', l)
    if m:
        outer = re.findall(r'inlined at "(.*?)", line (\d+)', m.group(3))
        f, ln = (outer[-1][0], int(outer[-1][1])) if outer else (m.group(1), int(m.group(2)))
        if f.endswith("swarm_kernels.cuh"): cur = ln
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): cnt[cur] += 1; seq.append(cur)
# phase markers: lines of the source that start with '    // ----' or contain 'PHASE:'
marks = [(1, "top")]
for k, l in enumerate(open(src), 1):
    m = re.search(r"// ---- ([^:.(]*)", l)
    if m: marks.append((k, m.group(1).strip()[:40]))
keys = [m[0] for m in marks]
b = collections.Counter()
for ln, c in cnt.items():
    if ln is None: b["?"] += c; continue
    b[marks[bisect.bisect_right(keys, ln) - 1]] += c
print("total", sum(cnt.values()))
for m in marks:
    if b[m]: print(f"{m[0]:5d} {m[1]:42s} {b[m]}")
if "-v" in sys.argv:
    for ln in sorted(k for k in cnt if k): print(ln, cnt[ln])
