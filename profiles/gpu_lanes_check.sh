mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_lanes.py -q -x 2>&1 | tail -40) > gpurun_out/lanes_tests.log 2>&1
(timeout 400 python profiles/run_lanes.py --tlogs 11,13 --c4) > gpurun_out/lanes.jsonl 2> gpurun_out/lanes.err
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:kmc_lanes -c 1 -s 1 -f -o gpurun_out/prof_lanes_r01_v1 python profiles/run_lanes.py --controls 1024 --kernels lanes) > gpurun_out/ncu_lanes.log 2>&1
tail -5 gpurun_out/lanes_tests.log; cat gpurun_out/lanes.jsonl; tail -3 gpurun_out/lanes.err; tail -3 gpurun_out/ncu_lanes.log
