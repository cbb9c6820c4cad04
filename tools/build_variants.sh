#!/bin/bash
# A/B builds of libswarm_b200.so into marl_llm_b200/lib/variants/<name>.so:  tools/build_variants.sh name "-DFLAG ..." [name flags]...
set -e
cd "$(dirname "$0")/.."
mkdir -p marl_llm_b200/lib/variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -shared -Xcompiler -fPIC $flags \
       -Xptxas -v -o marl_llm_b200/lib/variants/$name.so marl_llm_b200/csrc/swarm_abi.cu marl_llm_b200/csrc/rollout_abi.cu marl_llm_b200/csrc/policy_abi.cu 2>&1 \
       | grep -c "Used" | sed "s/^/[$name] kernels: /" &
done
wait
