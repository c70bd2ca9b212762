# Round 2, second GPU call: the tcgen05 candidate filter of daisy_topk_full (parity first, then the eval line), the whole
# GPU suite with the four un-gated units, funk-SVD after the cooperative-launch change.
set -x
timeout 600 python -m pytest tests/test_bpr_gpu.py -m gpu -q -x -k "topk_full" 2>&1 | tail -25 > gpurun_out/r02b_topk_tests.log
tail -5 gpurun_out/r02b_topk_tests.log
timeout 600 python bench.py --workload eval > gpurun_out/r02b_bench_eval.json 2> gpurun_out/r02b_bench_eval.err
cut -c1-1500 gpurun_out/r02b_bench_eval.json; tail -5 gpurun_out/r02b_bench_eval.err
DAISY_TOPK_TC=0 timeout 600 python bench.py --workload eval --steps 2 > gpurun_out/r02b_bench_eval_cuda_cores.json 2> /dev/null
cut -c1-300 gpurun_out/r02b_bench_eval_cuda_cores.json
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02b_gpu_tests.log
tail -8 gpurun_out/r02b_gpu_tests.log
timeout 300 python bench.py --workload config2 > gpurun_out/r02b_bench_config2.json 2> gpurun_out/r02b_bench_config2.err
cut -c1-400 gpurun_out/r02b_bench_config2.json
