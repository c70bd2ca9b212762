# Round 2, last 1-GPU call: the driver's commands on the final tree (bench at its flags, smoke, whole GPU suite), the
# default bench line with the lazy-decay fold timed beside the steps, and two build-time variants of the main kernel
# (5 / 6 resident blocks per SM: 48 / 40 registers).
set -x
timeout 600 python bench.py > gpurun_out/r02s_bench_n1_default.json 2> gpurun_out/r02s_bench_n1_default.err
show() { python -c "
import json; d=json.load(open('gpurun_out/r02s_bench_$1.json')); print('$1', round(d['ms_per_step'],4), 'value', round(d['value']/1e9,4), 'e2e', round(d['e2e']['value']/1e9,4), 'main', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step_frac'],4), 'mat', d.get('materialize_ms'))"; }
show n1_default
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02s_bench_n1_driver_flags.json 2>/dev/null; show n1_driver_flags
for v in main5 main6; do DAISY_LIB_VARIANT=$v timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02s_bench_$v.json 2>/dev/null; show $v; done
DAISY_SEG_WIN=16 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02s_bench_segwin16.json 2>/dev/null; show segwin16
DAISY_LIB_VARIANT=main5 timeout 300 python -m pytest tests/test_bpr_gpu.py -m gpu -q -x -k "not topk" 2>&1 | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 > gpurun_out/r02s_smoke.log; cat gpurun_out/r02s_smoke.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02s_gpu_tests.log; tail -3 gpurun_out/r02s_gpu_tests.log
