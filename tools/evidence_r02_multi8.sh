# Round 2, 8-GPU call (second: exclusive-row bypass on by default):  gpurun --gpus 8 --timeout 600 -- bash tools/evidence_r02_multi8.sh
# One-process-per-GPU parity (2, 4 and 8 ranks: step vs oracle, bit for bit vs the lockstep run; fit() vs the
# single-device ml-100k trajectory), then the bench line with its parity self-check.
set -x
timeout 420 python -m pytest tests/test_sharded_gpu.py -m gpu -q -k "multiprocess or fit" 2>&1 | tail -15 > gpurun_out/r02n_tests_n8.log
tail -6 gpurun_out/r02n_tests_n8.log
for N in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 50 --warmup 5 2> gpurun_out/r02n_bench_n${N}_default.err | grep '^{' > gpurun_out/r02n_bench_n${N}_default.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02n_bench_n${N}_default.json"))
    print("N=$N", round(d["ms_per_step"],4), "ms/step", round(d["value"]/1e9,4), "G/s e2e", round(d["e2e"]["value"]/1e9,4), d["config"].get("parity_selfcheck"))
    print("   phases", d.get("phase_ms(device,host)"))
    for r in d.get("phase_ms_by_rank") or []: print("   ", r)
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02n_bench_n${N}_default.err").read()[-1500:])
PY
done
