# Round 2, third GPU call (1 GPU): tcgen05 filter with reserved candidate slots, the whole GPU suite (routing, sharded
# evaluation in-process, Adam pin), ncu of the filter kernel.
set -x
timeout 600 python -m pytest tests/test_bpr_gpu.py -m gpu -q -x -k "topk_full or adam" 2>&1 | tail -15 > gpurun_out/r02c_topk_tests.log
tail -4 gpurun_out/r02c_topk_tests.log
timeout 600 python bench.py --workload eval > gpurun_out/r02c_bench_eval.json 2> gpurun_out/r02c_bench_eval.err
cut -c1-1400 gpurun_out/r02c_bench_eval.json; tail -5 gpurun_out/r02c_bench_eval.err
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02c_gpu_tests.log
tail -12 gpurun_out/r02c_gpu_tests.log
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_filter_tc -c 1 -o gpurun_out/r02c_filter_tc \
  python bench.py --workload eval --steps 1 > gpurun_out/r02c_ncu_eval.log 2>&1
tail -3 gpurun_out/r02c_ncu_eval.log
ls -la gpurun_out/*.ncu-rep
