# Round 1, final evidence set (one gpurun call, 1 GPU): tests, the driver's line, secondary lines, launch lists, ncu.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r01d_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01d_smoke.log 2>&1
python bench.py > gpurun_out/r01d_bench_n1_default.json 2> gpurun_out/r01d_bench_n1_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01d_bench_reference_arm.json 2>/dev/null
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --phases --trace > gpurun_out/r01d_bench_n1_phases_trace.json 2>/dev/null
for b in 4096 8192 16384 32768 65536; do python bench.py --workload config3 --batch $b --steps 400 --warmup 20 --no-cpu-baseline --epoch-api > gpurun_out/r01d_bench_config3_b$b.json 2>/dev/null; done
python bench.py --workload config3 --batch 131072 --steps 200 --warmup 20 --no-cpu-baseline --epoch-api > gpurun_out/r01d_bench_config3_b131072.json 2>/dev/null
python bench.py --workload config1 > gpurun_out/r01d_bench_config1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01d_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r01d_launches_b65536.csv python bench.py --workload config3 --batch 65536 --steps 4 --warmup 3 --no-cpu-baseline --epoch-api > gpurun_out/ncu_list_mid.log 2>&1
tail -3 gpurun_out/r01d_gpu_tests.log
cat gpurun_out/r01d_smoke.log | tail -2
for f in gpurun_out/r01d_bench_*.json; do echo $f; cut -c1-220 $f | tail -1; done
