#!/bin/bash
# Round-2 experiment 7: the 256-acceptor kernel (hop_wide.cu) at 3 / 4 / 5 CTAs per SM (register cap by __launch_bounds__, ring of
# 4 / 2 rows so that shared memory allows it).  C5, same box, all builds.  With arguments: the flag sets to build (first one not
# empty); WIDE_CONFIGS=C5-pruned measures the sparse sweep (-DSP_MIN_CTAS, -DSP_LOGK, -DSPC).
FLAGS=("" "-DWIDE_MIN_CTAS=3" "-DWIDE_MIN_CTAS=4 -DRING_D=2" "-DWIDE_MIN_CTAS=5 -DRING_D=2" "-DRING_D=2")
[ -n "$1" ] && FLAGS=("$@")
for flag in "${FLAGS[@]}"; do
  rm -f kmc_dn_b200/build/hop_wide*.o
  KMCB200_NVCC_FLAGS="$flag" python -m kmc_dn_b200.build > /dev/null
  echo "{\"nvcc_flags\": \"$flag\"}"
  cuobjdump -res-usage kmc_dn_b200/libkmcb200.so 2>/dev/null | grep -A1 "kmc_wide_kernelILi8ELi[34]ELb0ELb1" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | paste - -
  python profiles/run_configs.py --tag tmp --out-dir gpurun_out --only ${WIDE_CONFIGS:-C5} --no-cpu | cut -c1-220
done
