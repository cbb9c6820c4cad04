#!/bin/bash
# run bench.py once per A/B library under marl_llm_b200/lib/variants (device-resident number only): tools/ab.sh [names...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/variants
for v in "$@"; do
  SWARM_B200_LIB=$PWD/marl_llm_b200/lib/variants/$v.so python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/variants/$v.log 2>&1
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/variants/{v}.log").read().strip().splitlines()[-1])
    print(f"{v:12s} ms/step {d['ms_per_step']:.4f}  value {d['value']:.4g}  frac {d['roofline']['frac']:.3f}  clk {d['clocks']['sm_mhz']}")
except Exception as ex:
    print(v, "FAILED", ex)
PY
done
