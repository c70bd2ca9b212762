# Round 2, multi-GPU call:  gpurun --gpus N --timeout 900 -- 'bash tools/evidence_r02_multi.sh N'
#   parity first (one process per GPU: step vs oracle and vs the lockstep run bit for bit; fit() vs the single-device
#   ml-100k trajectory), then the bench line (with its own parity self-check), the schedule A/B and the stage sweep.
set -x
N=${1:-2}
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q -k "multiprocess or fit" 2>&1 | tail -15 > gpurun_out/r02m_tests_n$N.log
tail -6 gpurun_out/r02m_tests_n$N.log
run() {  # $1 = tag, rest = environment
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 50 --warmup 5 2> gpurun_out/r02m_bench_n${N}_${tag}.err | grep "^{" > gpurun_out/r02m_bench_n${N}_${tag}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02m_bench_n${N}_${tag}.json"))
    print("$tag", round(d["ms_per_step"],4), "ms/step", round(d["value"]/1e9,4), "G/s e2e", round(d["e2e"]["value"]/1e9,4), d["config"].get("parity_selfcheck"))
    print("   phases", d.get("phase_ms(device,host)"))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r02m_bench_n${N}_${tag}.err").read()[-1500:])
PY
}
run default DAISY_X=0
run sorted DAISY_SHARD_INTERLEAVE=0
run stages4 DAISY_MAIN_STAGES=4
timeout 60 ./tools/peer_a2a_bw $N > gpurun_out/r02m_peer_a2a_bw_${N}gpu.log 2>&1; tail -12 gpurun_out/r02m_peer_a2a_bw_${N}gpu.log
