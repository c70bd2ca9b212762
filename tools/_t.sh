python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload config1 > gpurun_out/r01d_bench_config1.json 2> gpurun_out/c1.err; tail -2 gpurun_out/c1.err; cut -c1-150 gpurun_out/r01d_bench_config1.json; python -c "
import json; d=json.loads(open('gpurun_out/r01d_bench_config1.json').read()); print(d['final'], d['device_sampler'], d['fit_wall_s'])"
python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-220
