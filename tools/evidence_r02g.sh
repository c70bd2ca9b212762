# Round 2, seventh GPU call (1 GPU): merged ref sort, SM partitioning (green contexts) and an occupancy cap of the main
# kernel, each against the shared-SM pipeline; parity tests under each.
set -x
b() {  # $1 = tag, rest = environment
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02g_bench_$tag.json 2> gpurun_out/r02g_bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02g_bench_$tag.json"))
    print("$tag", round(d["ms_per_step"],4), "ms/step", round(d["value"]/1e9,4), "G/s e2e", round(d["e2e"]["value"]/1e9,4), "main", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["whole_step_frac"],4))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r02g_bench_$tag.err").read()[-800:])
PY
}
b separate DAISY_MERGED_SORT=0
b merged DAISY_MERGED_SORT=1
b cap3 DAISY_MAIN_MAX_BLOCKS=3
b cap2 DAISY_MAIN_MAX_BLOCKS=2
b part8 DAISY_BOOK_SMS=8 DAISY_BOOK_SMS_REQUIRED=1
b part16 DAISY_BOOK_SMS=16 DAISY_BOOK_SMS_REQUIRED=1
b part24 DAISY_BOOK_SMS=24 DAISY_BOOK_SMS_REQUIRED=1
b part32 DAISY_BOOK_SMS=32 DAISY_BOOK_SMS_REQUIRED=1
timeout 300 python bench.py --no-cpu-baseline --phases > gpurun_out/r02g_bench_phases.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_phases.json')); print('phases', d.get('phase_ms'))"
DAISY_BOOK_SMS=16 timeout 300 python bench.py --no-cpu-baseline --trace --steps 12 > gpurun_out/r02g_bench_trace_part16.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_trace_part16.json')); print(d['trace_ms(book_begin,book_end,kernels_begin,kernels_end)'][4:9])"
timeout 300 python bench.py --no-cpu-baseline --trace --steps 12 > gpurun_out/r02g_bench_trace.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_trace.json')); print(d['trace_ms(book_begin,book_end,kernels_begin,kernels_end)'][4:9])"
DAISY_MERGED_SORT=1 timeout 600 python -m pytest tests/test_bpr_gpu.py tests/test_bprfm_gpu.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r02g_tests_merged.log; tail -2 gpurun_out/r02g_tests_merged.log
DAISY_BOOK_SMS=16 DAISY_BOOK_SMS_REQUIRED=1 timeout 600 python -m pytest tests/test_bpr_gpu.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r02g_tests_part16.log; tail -2 gpurun_out/r02g_tests_part16.log
