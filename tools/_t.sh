python -m pytest tests/test_bpr_gpu.py -m gpu -x -q 2>&1 | tail -3
for b in 16384 32768 65536 131072; do python bench.py --workload config3 --batch $b --steps 400 --warmup 20 --no-cpu-baseline --epoch-api --phases > gpurun_out/bench_c3m_b$b.log 2>&1; tail -1 gpurun_out/bench_c3m_b$b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['batch'], round(d['value']/1e6,1), round(d['ms_per_step']*1e3,1), 'e2e', round(d['e2e']['value']/1e6,1), 'host', d['host_enqueue_ms_per_step'], 'main', round(d['roofline']['kernel_ms']*1e3,1), d['gpu_launches']); print({k:v for k,v in d['phase_ms'].items() if v>0.004})"; done
DAISY_MID_MAX=0 python bench.py --workload config3 --batch 131072 --steps 400 --warmup 20 --no-cpu-baseline --epoch-api 2>/dev/null | tail -1 | cut -c1-200
