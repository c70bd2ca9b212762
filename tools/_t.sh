python -m pytest tests/test_gmf_gpu.py -m gpu -x -q 2>&1 | tail -4
python bench.py --workload gmf > gpurun_out/r01d_bench_gmf.json 2> gpurun_out/gmf.err; tail -2 gpurun_out/gmf.err; python -c "
import json; d=json.loads(open('gpurun_out/r01d_bench_gmf.json').read()); print(round(d['value']/1e6,2),'M samples/s', round(d['ms_per_step']*1e3,1),'us/step; e2e', round(d['e2e']['value']/1e6,2), 'cpu', round(d['cpu_baseline']['value']/1e6,3), 'launches', d['gpu_launches'], 'loss', d['final_loss'])"
