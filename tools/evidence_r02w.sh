# Round 2, closing call: k_sgns_main with the next row in flight (bench), whole GPU suite on the final tree.
set -x
timeout 300 python bench.py --workload sgns --no-cpu-baseline > gpurun_out/r02w_bench_sgns.json 2> gpurun_out/r02w_bench_sgns.err
python -c "
import json; d=json.load(open('gpurun_out/r02w_bench_sgns.json')); print('sgns', round(d['value']), d['unit'], round(d['ms_per_step'],4), 'ms/step e2e', round(d['e2e']['value']))"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02w_gpu_tests.log; tail -2 gpurun_out/r02w_gpu_tests.log
