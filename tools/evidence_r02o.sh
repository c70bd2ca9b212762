# Round 2, 1-GPU call: tensor-core filter with software-pipelined TMEM loads (8 and 16 epilogue warps).
set -x
timeout 600 python -m pytest tests/test_bpr_gpu.py -m gpu -q -x -k "topk_full" 2>&1 | tail -3 > gpurun_out/r02o_topk_tests.log; tail -2 gpurun_out/r02o_topk_tests.log
DAISY_LIB_VARIANT=epi16 timeout 600 python -m pytest tests/test_bpr_gpu.py -m gpu -q -x -k "topk_full" 2>&1 | tail -3 > gpurun_out/r02o_topk_tests_epi16.log; tail -2 gpurun_out/r02o_topk_tests_epi16.log
for v in default epi16; do
  if [ $v = default ]; then unset DAISY_LIB_VARIANT; else export DAISY_LIB_VARIANT=$v; fi
  DAISY_TC_STATS=1 timeout 300 python bench.py --workload eval --no-cpu-baseline > gpurun_out/r02o_bench_eval_$v.json 2> gpurun_out/r02o_bench_eval_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r02o_bench_eval_$v.json')); print('$v', round(d['ms_per_step'],2), 'ms', round(d['value']), 'users/s filter', round(d['roofline']['kernel_ms'],2), 'rescore', round(d['roofline']['rescore_kernel_ms'],2), 'frac', round(d['roofline']['frac'],3))"
  grep "k_filter_tc. kernel" gpurun_out/r02o_bench_eval_$v.err | tail -1 | cut -c1-330
done
