# Round 2, fifth GPU call (1 GPU): where the tensor-core filter's time goes (per-role wait accounting), 16 epilogue warps,
# step timeline of config 4, the new sharded-path tests that run on one device.
set -x
DAISY_TC_STATS=1 timeout 300 python bench.py --workload eval --steps 1 > gpurun_out/r02e_eval_stats.json 2> gpurun_out/r02e_eval_stats.err
grep k_filter_tc gpurun_out/r02e_eval_stats.err | tail -2
DAISY_LIB_VARIANT=epi16 DAISY_TC_STATS=1 timeout 300 python bench.py --workload eval --steps 1 > gpurun_out/r02e_eval_stats_epi16.json 2> gpurun_out/r02e_eval_stats_epi16.err
grep k_filter_tc gpurun_out/r02e_eval_stats_epi16.err | tail -2
DAISY_LIB_VARIANT=epi16 timeout 300 python bench.py --workload eval > gpurun_out/r02e_bench_eval_epi16.json 2>/dev/null
cut -c1-120 gpurun_out/r02e_bench_eval_epi16.json; python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_eval_epi16.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['rescore_kernel_ms'])"
DAISY_LIB_VARIANT=epi16 timeout 600 python -m pytest tests/test_bpr_gpu.py -m gpu -q -x -k "topk_full" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_sharded_gpu.py tests/test_bprfm_gpu.py tests/test_neumf_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r02e_sharded_tests.log
tail -6 gpurun_out/r02e_sharded_tests.log
timeout 300 python bench.py --no-cpu-baseline --trace --steps 12 > gpurun_out/r02e_bench_trace.json 2> gpurun_out/r02e_bench_trace.err
tail -c 2500 gpurun_out/r02e_bench_trace.json
