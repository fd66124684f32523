#!/bin/bash
# Round-2 experiment 6: what the reference-exact Boltzmann factor (kmc_device.cuh boltz: 8 instructions per evaluated pair
# instead of 3) costs on C3 (thread-per-trajectory kernel) and C5 (wide kernel).  Same box, both builds.
for flag in "-DKMCB200_FAST_BOLTZ" ""; do
  rm -f kmc_dn_b200/build/*.o
  KMCB200_NVCC_FLAGS="$flag" python -m kmc_dn_b200.build > /dev/null
  echo "{\"nvcc_flags\": \"$flag\"}"
  python profiles/run_lanes.py --controls 16384 --hops 100000 --kernels lanes | cut -c1-150
  python profiles/run_configs.py --tag tmp --only C5,C2,C4 --no-cpu | cut -c1-220
done
