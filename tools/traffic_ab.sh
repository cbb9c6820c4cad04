cd "$(dirname "$0")/.."
for v in "$@"; do
  SWARM_B200_LIB=$PWD/marl_llm_b200/lib/variants/$v.so ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_step -s 8 -c 2 --csv --log-file gpurun_out/traffic_$v.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
  python - "$v" <<'PY'
import csv,sys
v=sys.argv[1]
rows=[r for r in csv.reader(l for l in open(f"gpurun_out/traffic_{v}.csv") if not l.startswith("=="))]
h=rows[0]; ni=h.index("Metric Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit"); ki=h.index("Kernel Name")
tot=0
for r in rows[1:]:
    if r[ni].startswith("dram"):
        m={"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}[r[ui]]; tot+=float(r[vi].replace(",",""))*m
print(v, "dram GB per step", round(tot/1e9,3))
PY
  SWARM_B200_LIB=$PWD/marl_llm_b200/lib/variants/$v.so python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   ms', round(d['ms_per_step'],4))"
done
