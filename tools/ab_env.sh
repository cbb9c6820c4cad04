#!/bin/bash
# A/B of one environment variable on one box: tools/ab_env.sh "bench args" VAR v1 v2 ...   (device-resident numbers, both regimes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/variants
args=$1; var=$2; shift 2
for v in "$@"; do
  env $var=$v python bench.py $args --no-e2e --no-cpu-baseline --no-extras > gpurun_out/variants/$var.$v.log 2>&1
  python - "$var.$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/variants/{v}.log").read().strip().splitlines()[-1])
    print(f"{v:18s} " + "  ".join(f"{k} {r['ms_per_step']:.4f} ms (frac {r['roofline']['frac']:.3f})" for k, r in d["regimes"].items()) + f"  clk {d['clocks']['sm_mhz']}")
except Exception as ex:
    print(v, "FAILED", ex)
PY
done
