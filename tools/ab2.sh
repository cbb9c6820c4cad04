#!/bin/bash
# A/B of library variants on one box: tools/ab2.sh "bench args" name1 name2 ...   (device-resident number only, both regimes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/variants
args=$1; shift
for v in "$@"; do
  lib=$PWD/marl_llm_b200/lib/variants/$v.so
  [ "$v" = base ] && lib=$PWD/marl_llm_b200/lib/libswarm_b200.so
  SWARM_B200_LIB=$lib python bench.py $args --no-e2e --no-cpu-baseline --no-extras > gpurun_out/variants/$v.log 2>&1
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/variants/{v}.log").read().strip().splitlines()[-1])
    print(f"{v:14s} " + "  ".join(f"{k} {r['ms_per_step']:.4f} ms (frac {r['roofline']['frac']:.3f})" for k, r in d["regimes"].items()) + f"  clk {d['clocks']['sm_mhz']}")
except Exception as ex:
    print(v, "FAILED", ex)
PY
done
