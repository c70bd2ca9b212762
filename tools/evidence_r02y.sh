set -x
DAISY_SVDPP_STATS=1 timeout 300 python bench.py --workload svdpp --no-cpu-baseline > gpurun_out/r02y_bench_svdpp.json 2> gpurun_out/r02y_bench_svdpp.err
grep daisy_svdpp_fit gpurun_out/r02y_bench_svdpp.err | tail -1
