# Round 2, fourth GPU call (1 GPU): tcgen05 filter with 8 epilogue warps + max-tree scan; L2 access-policy window experiment
# on config 4 (hot items as a prefix of the item table = what a popularity remap would produce).
set -x
timeout 600 python -m pytest tests/test_bpr_gpu.py -m gpu -q -x -k "topk_full" 2>&1 | tail -6 > gpurun_out/r02d_topk_tests.log
tail -3 gpurun_out/r02d_topk_tests.log
timeout 600 python bench.py --workload eval > gpurun_out/r02d_bench_eval.json 2> gpurun_out/r02d_bench_eval.err
cut -c1-1200 gpurun_out/r02d_bench_eval.json; tail -5 gpurun_out/r02d_bench_eval.err
for v in "permuted:" "prefix:--hot-prefix" "prefix_win32k:--hot-prefix --l2-window --l2-window-rows 32768" \
         "prefix_win128k:--hot-prefix --l2-window --l2-window-rows 131072" "permuted_winall:--l2-window"; do
  tag=${v%%:*}; flags=${v#*:}
  timeout 300 python bench.py --no-cpu-baseline $flags > gpurun_out/r02d_bench_n1_$tag.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/r02d_bench_n1_$tag.json"))
print("$tag", round(d["ms_per_step"],4), "ms/step", round(d["value"]/1e9,4), "G/s main", round(d["roofline"]["kernel_ms"],4))
PY
done
timeout 600 ncu --metrics lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  --kernel-name regex:'k_bpr_main_tma|k_seg_all' --launch-skip 20 -c 4 --csv --log-file gpurun_out/r02d_ncu_prefix_win128k.csv \
  python bench.py --no-cpu-baseline --hot-prefix --l2-window --l2-window-rows 131072 --steps 10 > /dev/null 2>&1
timeout 600 ncu --metrics lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  --kernel-name regex:'k_bpr_main_tma|k_seg_all' --launch-skip 20 -c 4 --csv --log-file gpurun_out/r02d_ncu_prefix_nowin.csv \
  python bench.py --no-cpu-baseline --hot-prefix --steps 10 > /dev/null 2>&1
tail -4 gpurun_out/r02d_ncu_prefix_win128k.csv | cut -c1-300; tail -4 gpurun_out/r02d_ncu_prefix_nowin.csv | cut -c1-300
