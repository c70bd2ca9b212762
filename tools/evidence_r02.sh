# Round 2, first GPU calls: what round 1 could not measure after its GPU budget ended.
#   1 GPU :  gpurun --timeout 1500 -- 'bash tools/evidence_r02.sh one'   (after: bash tools/build_variants.sh)
#   N GPUs:  gpurun --gpus 8 --timeout 400 -- 'bash tools/evidence_r02.sh many 8'     (then 4, 2)
set -x
mode=${1:-one}
if [ "$mode" = one ]; then
  python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_gpu_tests.log
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
  python bench.py > gpurun_out/r02_bench_n1_default.json 2> gpurun_out/r02_bench_n1_default.err
  tail -3 gpurun_out/r02_gpu_tests.log; tail -1 gpurun_out/r02_smoke.log; cut -c1-300 gpurun_out/r02_bench_n1_default.json
  # k_seg_all variants (lead 3).  Build them HERE before the call: bash tools/build_variants.sh
  python bench.py --no-cpu-baseline --phases > gpurun_out/r02_bench_n1_phases.json 2>/dev/null
  cut -c1-200 gpurun_out/r02_bench_n1_phases.json
  for v in segmb6 segmb8 segpf segpf8; do
    if [ -f recommend_lib_b200/libdaisy_b200_$v.so ]; then
      DAISY_LIB_VARIANT=$v python -m pytest tests/test_bpr_gpu.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_gpu_tests_$v.log
      DAISY_LIB_VARIANT=$v python bench.py --no-cpu-baseline --phases > gpurun_out/r02_bench_n1_$v.json 2>/dev/null
      tail -1 gpurun_out/r02_gpu_tests_$v.log; cut -c1-200 gpurun_out/r02_bench_n1_$v.json
    fi
  done
  # experimental kernels that have never run on a GPU (last: a crash here must not cost the evidence above)
  DAISY_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_bprfm_bn_gpu.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_experimental_fmbn.log
  tail -5 gpurun_out/r02_experimental_fmbn.log
  DAISY_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_sgns_gpu.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_experimental_sgns.log
  tail -5 gpurun_out/r02_experimental_sgns.log
  DAISY_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_neumf_gpu.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_experimental_neumf.log
  tail -5 gpurun_out/r02_experimental_neumf.log
  DAISY_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_svdpp_gpu.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_experimental_svdpp.log
  tail -5 gpurun_out/r02_experimental_svdpp.log
  for w in bprfm_bn sgns neumf svdpp; do timeout 300 python bench.py --workload $w > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; cut -c1-220 gpurun_out/r02_bench_$w.json; done
  # SVD++: block size of the one-block kernel (barrier cost against rows in flight)
  DAISY_SVDPP_HOT=0 timeout 300 python bench.py --workload svdpp --steps 1 > gpurun_out/r02_bench_svdpp_nohot.json 2>/dev/null; cut -c1-120 gpurun_out/r02_bench_svdpp_nohot.json
  for t in 128 256 512; do DAISY_SVDPP_THREADS=$t timeout 300 python bench.py --workload svdpp --steps 1 > gpurun_out/r02_bench_svdpp_t$t.json 2>/dev/null; cut -c1-120 gpurun_out/r02_bench_svdpp_t$t.json; done
else
  N=${2:-8}
  run() {  # $1 = tag, rest = environment
    tag=$1; shift
    env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02_bench_n${N}_${tag}.json 2> gpurun_out/r02_bench_n${N}_${tag}.err
    cut -c1-200 gpurun_out/r02_bench_n${N}_${tag}.json
  }
  run interleaved DAISY_SHARD_INTERLEAVE=$N     # the default since r01e: chunks dealt round-robin over the owners
  run sorted DAISY_SHARD_INTERLEAVE=0           # the schedule of every r01b-r01d multi-GPU line
  timeout 60 ./tools/peer_a2a_bw $N > gpurun_out/r02_peer_a2a_bw_${N}gpu.log 2>&1; cat gpurun_out/r02_peer_a2a_bw_${N}gpu.log
fi
