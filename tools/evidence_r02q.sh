# Round 2, 2-GPU call: shared rows of the owner side through per-row slots (one launch, no searches) against the passes.
set -x
N=${1:-2}
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q -x -k "bypass or lockstep" 2>&1 | tail -8 > gpurun_out/r02q_tests_lockstep.log
tail -3 gpurun_out/r02q_tests_lockstep.log
DAISY_SHARD_BYPASS=1 timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q -k "multiprocess or fit" 2>&1 | tail -8 > gpurun_out/r02q_tests_bypass_n$N.log
tail -3 gpurun_out/r02q_tests_bypass_n$N.log
run() {  # $1 = tag, rest = environment
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 50 --warmup 5 2> gpurun_out/r02q_bench_n${N}_${tag}.err | grep "^{" > gpurun_out/r02q_bench_n${N}_${tag}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02q_bench_n${N}_${tag}.json"))
    print("$tag", round(d["ms_per_step"],4), "ms/step", round(d["value"]/1e9,4), "G/s e2e", round(d["e2e"]["value"]/1e9,4), d["config"].get("parity_selfcheck"))
    print("   phases", d.get("phase_ms(device,host)"))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r02q_bench_n${N}_${tag}.err").read()[-1500:])
PY
}
run bypass DAISY_SHARD_BYPASS=1
b() {  # 1-GPU side experiments
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02q_bench_$tag.json 2> gpurun_out/r02q_bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02q_bench_$tag.json"))
    print("$tag", round(d["ms_per_step"],4), "ms/step", round(d["value"]/1e9,4), "G/s e2e", round(d["e2e"]["value"]/1e9,4), "main", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["whole_step_frac"],4))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r02q_bench_$tag.err").read()[-800:])
PY
}



run passes DAISY_OWNER_SHARED_PASSES=1
