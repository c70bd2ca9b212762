# Round 2, closing 1-GPU call: SVD++ with resident rows off by default (tests + bench line), smoke on the final tree.
set -x
timeout 600 python -m pytest tests/test_svdpp_gpu.py tests/test_sgns_gpu.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r02t_tests.log; tail -2 gpurun_out/r02t_tests.log
timeout 600 python bench.py --workload svdpp > gpurun_out/r02t_bench_svdpp.json 2> gpurun_out/r02t_bench_svdpp.err
python -c "
import json; d=json.load(open('gpurun_out/r02t_bench_svdpp.json')); print('svdpp', round(d['value']), d['unit'], d['cpu_baseline']['value'])"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
