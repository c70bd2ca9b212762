# Round 2, 1-GPU call: copy stream for host triples (e2e), funk-SVD after the metadata change, fresh default bench line,
# ncu capture of the current tensor-core filter.
set -x
timeout 900 python -m pytest tests/test_bpr_gpu.py tests/test_mf_gpu.py tests/test_gmf_gpu.py tests/test_bprfm_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02l_tests.log; tail -2 gpurun_out/r02l_tests.log
timeout 600 python bench.py > gpurun_out/r02l_bench_n1_default.json 2> gpurun_out/r02l_bench_n1_default.err
python -c "
import json; d=json.load(open('gpurun_out/r02l_bench_n1_default.json')); print('config4', round(d['ms_per_step'],4), 'value', round(d['value']/1e9,4), 'e2e', round(d['e2e']['value']/1e9,4), round(d['e2e']['ms_per_step'],4), 'main', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['whole_step_frac'],4), d['cpu_baseline'])"
DAISY_COPY_STREAM=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02l_bench_n1_nocopystream.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02l_bench_n1_nocopystream.json')); print('no copy stream: value', round(d['value']/1e9,4), 'e2e', round(d['e2e']['value']/1e9,4), round(d['e2e']['ms_per_step'],4))"
timeout 600 python bench.py --epoch-api --no-cpu-baseline > gpurun_out/r02l_bench_n1_epoch_api.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02l_bench_n1_epoch_api.json')); print('epoch api: value', round(d['value']/1e9,4), 'e2e', round(d['e2e']['value']/1e9,4), round(d['e2e']['ms_per_step'],4))"
DAISY_MF_STATS=1 timeout 600 python bench.py --workload config2 > gpurun_out/r02l_bench_config2.json 2> gpurun_out/r02l_bench_config2.err
python -c "
import json; d=json.load(open('gpurun_out/r02l_bench_config2.json')); print('config2', round(d['value']/1e6,2), 'M ratings/s', round(d['ms_per_step'],2), 'ms/epoch')"
grep daisy_mf_fit gpurun_out/r02l_bench_config2.err | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_filter_tc -c 1 -o gpurun_out/r02l_filter_tc python bench.py --workload eval --steps 1 --no-cpu-baseline > gpurun_out/r02l_ncu_eval.log 2>&1
ncu -i gpurun_out/r02l_filter_tc.ncu-rep --page raw --csv > gpurun_out/r02l_filter_tc_raw.csv 2>/dev/null
rm -f gpurun_out/r02l_filter_tc.ncu-rep
