# Round 2, last 8-GPU call: parity at 8 ranks + fit, bench at 8 and 4 GPUs with the slot-based owner side; one A/B with the
# bookkeeping stream at low priority (the main kernel is NVLink-bound there and leaves SM time).
set -x
timeout 420 python -m pytest tests/test_sharded_gpu.py -m gpu -q -k "multiprocess and 8 or fit" 2>&1 | tail -15 > gpurun_out/r02r_tests_n8.log
tail -4 gpurun_out/r02r_tests_n8.log
run() {  # $1 = N, $2 = tag, rest = environment
  N=$1; tag=$2; shift; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 50 --warmup 5 2> gpurun_out/r02r_bench_n${N}_${tag}.err | grep '^{' > gpurun_out/r02r_bench_n${N}_${tag}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02r_bench_n${N}_${tag}.json"))
    print("N=$N $tag", round(d["ms_per_step"],4), "ms/step", round(d["value"]/1e9,4), "G/s e2e", round(d["e2e"]["value"]/1e9,4), d["config"].get("parity_selfcheck"))
    print("   phases", d.get("phase_ms(device,host)"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02r_bench_n${N}_${tag}.err").read()[-1500:])
PY
}
run 8 default DAISY_X=0
run 8 lowprio DAISY_BOOK_LOW_PRIORITY=1
run 4 default DAISY_X=0
