# lanes kernel on a B200: parity tests, variant timings, ncu capture (outputs under gpurun_out/)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_lanes.py -q -x 2>&1 | tail -40) > gpurun_out/lanes_tests.log 2>&1
(timeout 600 python profiles/run_lanes.py --kernels lanes --tlogs 13,14 --c4 --crossover 65536,131072) > gpurun_out/lanes.jsonl 2> gpurun_out/lanes.err
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:kmc_lanes -c 1 -s 1 -f -o gpurun_out/prof_lanes_r01_vN python profiles/run_lanes.py --controls 16384 --kernels lanes) > gpurun_out/ncu_lanes.log 2>&1
tail -5 gpurun_out/lanes_tests.log; python - <<'PY'
import json
for l in open('gpurun_out/lanes.jsonl'):
    r=json.loads(l); print(r['workload'], r['kernel'], r['members'], r.get('ltab_log'), '%.3e'%r['hops_per_s'], '%.1f ms'%r['ms_per_step'])
PY
tail -3 gpurun_out/lanes.err; tail -3 gpurun_out/ncu_lanes.log
(timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_lanes.jsonl 2> gpurun_out/bench_lanes.err); tail -c 3000 gpurun_out/bench_lanes.jsonl; tail -3 gpurun_out/bench_lanes.err
