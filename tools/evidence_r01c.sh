set -x
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r01c_gpu_tests.log
python bench.py > gpurun_out/r01c_bench_n1_default.json 2> gpurun_out/r01c_bench_n1_default.err
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --phases --trace > gpurun_out/r01c_bench_n1_phases_trace.json 2>/dev/null
for b in 4096 8192 65536; do python bench.py --workload config3 --batch $b --steps 400 --warmup 20 --no-cpu-baseline --epoch-api > gpurun_out/r01c_bench_config3_b$b.json 2>/dev/null; done
DAISY_SMALL_MAX=0 python bench.py --workload config3 --batch 4096 --steps 400 --warmup 20 --no-cpu-baseline --epoch-api > gpurun_out/r01c_bench_config3_b4096_general_path.json 2>/dev/null
python bench.py --workload config1 > gpurun_out/r01c_bench_config1.json 2> gpurun_out/r01c_bench_config1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01c_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bpr_main|k_seg_all" -s 6 -c 4 -o gpurun_out/r01c_prof python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r01c_launches_small.csv python bench.py --workload config3 --batch 4096 --steps 4 --warmup 3 --no-cpu-baseline --epoch-api > gpurun_out/ncu_list_small.log 2>&1
tail -3 gpurun_out/r01c_gpu_tests.log
cut -c1-300 gpurun_out/r01c_bench_n1_default.json
cut -c1-200 gpurun_out/r01c_bench_config1.json
tail -3 gpurun_out/r01c_bench_config1.err
