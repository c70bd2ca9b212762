# Round 2: SVD++ with the history rows kept in registers between the gather and the update (tests + bench line).
set -x
timeout 600 python -m pytest tests/test_svdpp_gpu.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r02x_tests.log; tail -2 gpurun_out/r02x_tests.log
timeout 600 python bench.py --workload svdpp > gpurun_out/r02x_bench_svdpp.json 2> gpurun_out/r02x_bench_svdpp.err
python -c "
import json; d=json.load(open('gpurun_out/r02x_bench_svdpp.json')); print('svdpp', round(d['value']), d['unit'], 'cpu', round(d['cpu_baseline']['value']))"
