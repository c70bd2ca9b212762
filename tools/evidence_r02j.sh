# Round 2, 1-GPU call: funk-SVD batched groups (parity, reproducibility, config-2 bench A/B), single-launch owner pass
# of the sharded step (lockstep bit-identity), round-2 ncu launch list and full capture of the two table kernels.
set -x
timeout 900 python -m pytest tests/test_mf_gpu.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02j_tests_mf.log; tail -3 gpurun_out/r02j_tests_mf.log
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q -x -k "bypass or lockstep" 2>&1 | tail -6 > gpurun_out/r02j_tests_shard.log; tail -3 gpurun_out/r02j_tests_shard.log
for m in 1 0; do
  DAISY_MF_BATCH=$m DAISY_MF_STATS=1 timeout 600 python bench.py --workload config2 > gpurun_out/r02j_bench_config2_batch$m.json 2> gpurun_out/r02j_bench_config2_batch$m.err
  python -c "
import json; d=json.load(open('gpurun_out/r02j_bench_config2_batch$m.json')); print('config2 batch=$m', round(d['value']/1e6,2), 'M ratings/s', round(d['ms_per_step'],2), 'ms/epoch', d.get('cpu_baseline'))"
  grep daisy_mf_fit gpurun_out/r02j_bench_config2_batch$m.err | tail -1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02j_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bpr_main|k_seg_all" -s 6 -c 2 -o gpurun_out/r02j_prof python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_ncu_full.log 2>&1
ncu -i gpurun_out/r02j_prof.ncu-rep --page raw --csv > gpurun_out/r02j_prof_raw.csv 2>/dev/null
rm -f gpurun_out/r02j_prof.ncu-rep
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02j_gpu_tests.log; tail -3 gpurun_out/r02j_gpu_tests.log
