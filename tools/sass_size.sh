#!/bin/bash
# SASS instruction count of the headline kernels of a library: tools/sass_size.sh lib.so   (profiling aid)
for pat in 'k_stepIfLb1ELb0ELi128ELi1ELi0E' 'k_stepIfLb0ELb0ELi128ELi2ELi2E' 'k_stepIfLb1ELb0ELi1024ELi0ELi2E'; do
  k=$(cuobjdump -res-usage "$1" 2>/dev/null | grep -o "_ZN5swarm6${pat}[A-Za-z0-9_]*" | sort -u | head -1)
  n=$(cuobjdump -sass -fun $k "$1" 2>/dev/null | grep -cE '^\s+/\*[0-9a-f]{4,}\*/')
  echo "$k $n instr $((n*16/1024)) KB"
done
