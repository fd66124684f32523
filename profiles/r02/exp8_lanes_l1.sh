#!/bin/bash
# Round-2 experiment 8: shared memory per CTA of the thread-per-trajectory kernel against the carve-out steps (7 CTAs x 24.9 KB
# = 196 KB carve-out, 56 KB of L1; <= 23.4 KB per CTA = 164 KB carve-out, 88 KB of L1).  LANES_RMAX = runs of a warp whose
# parameters live in shared memory (176 B each); LANES_WARPS x LANES_MIN_CTAS = 28 warps per SM in CTAs of 4 / 7 / 14 / 28 warps (the
# layout tables, 10 KB, are per CTA).  C3 1 048 576 x 1e5, same box, all builds.
FLAGS=("-DLANES_RMAX=8" "-DLANES_RMAX=6" "-DLANES_RMAX=4")
[ -n "$1" ] && FLAGS=("$@")
for flag in "${FLAGS[@]}"; do
  rm -f kmc_dn_b200/build/hop_lanes*.o
  KMCB200_NVCC_FLAGS="$flag" python -m kmc_dn_b200.build > /dev/null
  echo "{\"nvcc_flags\": \"$flag\"}"
  python profiles/run_lanes.py --controls 16384 --hops 100000 --kernels lanes | cut -c1-150
  python profiles/run_lanes.py --controls 1024 --hops 1000000 --kernels lanes | cut -c1-150
  [ -n "$EXP8_MORE" ] && python profiles/run_lanes.py --controls 512 --hops 100000 --kernels lanes | cut -c1-150
  [ -n "$EXP8_MORE" ] && python profiles/run_lanes.py --c4 --kernels lanes | tail -1 | cut -c1-150
done
