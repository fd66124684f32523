# round-1 final validation on one B200: GPU test suite, smoke, bench (both arms), ncu launch list + full capture of the
# production kernel, sanitizers over the thread-per-trajectory kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/final_tests.log 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/final_smoke.log 2>&1
(timeout 600 python bench.py) > gpurun_out/bench_n1.jsonl 2> gpurun_out/bench_n1.err
(timeout 600 python bench.py --impl reference) > gpurun_out/bench_n1_reference.jsonl 2> gpurun_out/bench_ref.err
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_bench_lanes.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline) > gpurun_out/bench_under_ncu.log 2>&1
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:kmc_lanes -c 1 -s 1 -f -o gpurun_out/prof_lanes_r01_final python profiles/run_lanes.py --kernels lanes) > gpurun_out/ncu_lanes.log 2>&1
(timeout 300 compute-sanitizer --tool memcheck python profiles/sanitizer_run.py --lanes-only) > gpurun_out/sanitizer_memcheck_lanes_r01.log 2>&1
(timeout 300 compute-sanitizer --tool racecheck python profiles/sanitizer_run.py --lanes-only) > gpurun_out/sanitizer_racecheck_lanes_r01.log 2>&1
(timeout 400 python profiles/run_lanes.py --c4 --crossover 65536,131072,262144) > gpurun_out/lanes.jsonl 2> gpurun_out/lanes.err
tail -4 gpurun_out/final_tests.log; tail -2 gpurun_out/final_smoke.log; cut -c1-400 gpurun_out/bench_n1.jsonl; cut -c1-400 gpurun_out/bench_n1_reference.jsonl; tail -n 2 gpurun_out/sanitizer_memcheck_lanes_r01.log; tail -n 2 gpurun_out/sanitizer_racecheck_lanes_r01.log
