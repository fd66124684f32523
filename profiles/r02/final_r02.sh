#!/bin/bash
# Round-2 final measurement pass on one B200 (gpurun --timeout 2400 -- bash profiles/r02/final_r02.sh).  Everything lands in gpurun_out/.
set -x
python bench.py --impl reference > gpurun_out/bench_r02_n1_reference.jsonl 2> gpurun_out/bench_r02_ref.err
python bench.py > gpurun_out/bench_r02_n1.jsonl 2> gpurun_out/bench_r02_n1.err
python bench.py --workload c5 > gpurun_out/bench_r02_c5_n1.jsonl 2> gpurun_out/bench_r02_c5_n1.err
# launch list of the bench command (numbers printed under ncu are not bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-long --ps-sims 0 > gpurun_out/bench_under_ncu.log 2>&1
# one full capture of the dominant kernel on the headline configuration
ncu --set full --clock-control none --import-source on -k regex:kmc_lanes -s 1 -c 1 -f -o gpurun_out/prof_lanes_r02_final \
    python profiles/run_lanes.py --controls 16384 --hops 100000 --kernels lanes > gpurun_out/ncu_lanes.log 2>&1
python profiles/run_configs.py --tag r02 --out-dir gpurun_out > gpurun_out/configs_r02.jsonl 2> gpurun_out/configs_r02.err
tail -c 600 gpurun_out/bench_r02_n1.jsonl
