N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --phases --trace > gpurun_out/r01c_bench_n${N}_phases.json 2> gpurun_out/r01c_bench_n$N.err
grep '^{' gpurun_out/r01c_bench_n${N}_phases.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']/1e9,3), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']/1e9,3), d['phase_ms(device,host)'])
for r in d.get('phase_ms_by_rank',[]): print(r)"
grep "^rank" gpurun_out/r01c_bench_n${N}_phases.json | cut -c1-400
tail -2 gpurun_out/r01c_bench_n$N.err
