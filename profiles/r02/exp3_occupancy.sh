#!/bin/bash
# Round-2 experiment 3: resident CTAs per SM of the thread-per-trajectory kernel (register budget 64 / 72 / 80) on C3.
for c in 8 7 6; do
  rm -f kmc_dn_b200/build/hop_lanes.o
  KMCB200_NVCC_FLAGS="-DLANES_MIN_CTAS=$c" python -m kmc_dn_b200.build > /dev/null
  echo "{\"min_ctas\": $c}"
  python profiles/run_lanes.py --controls 16384 --hops 100000 --kernels lanes
done
rm -f kmc_dn_b200/build/hop_lanes.o; python -m kmc_dn_b200.build > /dev/null
