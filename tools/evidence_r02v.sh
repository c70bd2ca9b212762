# Round 2: SGNS with length-proportional row slices and batch-norm BPR-FM with sliced column / row reductions: parity
# tests and bench lines (first versions: 2.36 ms and 0.43 ms per step, profiles/r02a_bench_{sgns,bprfm_bn}.json).
set -x
timeout 600 python -m pytest tests/test_sgns_gpu.py tests/test_bprfm_bn_gpu.py tests/test_bprfm_gpu.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r02v_tests.log; tail -2 gpurun_out/r02v_tests.log
for wl in sgns bprfm_bn; do
  timeout 300 python bench.py --workload $wl > gpurun_out/r02v_bench_$wl.json 2> gpurun_out/r02v_bench_$wl.err
  python -c "
import json; d=json.load(open('gpurun_out/r02v_bench_$wl.json')); print('$wl', round(d['value']), d['unit'], round(d['ms_per_step'],4), 'ms/step e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value']), 'launches', d['gpu_launches'])"
done
