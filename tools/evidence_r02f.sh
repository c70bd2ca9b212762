# Round 2, sixth GPU call (1 GPU): three bookkeeping sets (the side stream one step further ahead); floors of the
# tensor-core filter (epilogue that only reads TMEM / does nothing); whole GPU suite.
set -x
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02f_bench_n1.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02f_bench_n1.json')); print('config4', d['ms_per_step'], d['value']/1e9, d['e2e']['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['whole_step_frac'])"
timeout 300 python bench.py --no-cpu-baseline --trace --steps 12 > gpurun_out/r02f_bench_trace.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02f_bench_trace.json')); print(d['trace_ms(book_begin,book_end,kernels_begin,kernels_end)'][4:9])"
for m in 0 1 2; do
  DAISY_TC_DEBUG=$m DAISY_TC_STATS=1 timeout 300 python bench.py --workload eval --steps 1 > /dev/null 2> gpurun_out/r02f_eval_dbg$m.err
  grep k_filter_tc gpurun_out/r02f_eval_dbg$m.err | tail -2
  DAISY_LIB_VARIANT=epi16 DAISY_TC_DEBUG=$m DAISY_TC_STATS=1 timeout 300 python bench.py --workload eval --steps 1 > /dev/null 2> gpurun_out/r02f_eval_dbg${m}_epi16.err
  grep k_filter_tc gpurun_out/r02f_eval_dbg${m}_epi16.err | tail -2
done
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02f_gpu_tests.log
tail -5 gpurun_out/r02f_gpu_tests.log
