# Round 2: SVD++ with the next rating's list loads issued ahead of the updates (tests, bench, phase cycles).
set -x
timeout 300 python -m pytest tests/test_svdpp_gpu.py -m gpu -q 2>&1 | tail -2 > gpurun_out/r02z_tests.log; tail -1 gpurun_out/r02z_tests.log
timeout 300 python bench.py --workload svdpp --no-cpu-baseline > gpurun_out/r02z_bench_svdpp.json 2> gpurun_out/r02z_bench_svdpp.err
DAISY_SVDPP_STATS=1 timeout 300 python bench.py --workload svdpp --no-cpu-baseline > /dev/null 2> gpurun_out/r02z_bench_svdpp_stats.err
grep daisy_svdpp_fit gpurun_out/r02z_bench_svdpp_stats.err | tail -1
python -c "
import json; d=json.load(open('gpurun_out/r02z_bench_svdpp.json')); print('svdpp', round(d['value']), d['unit'])"
