set -x
r() { tag=$1; shift
env "$@" DAISY_MF_STATS=1 timeout 600 python bench.py --workload config2 --no-cpu-baseline > gpurun_out/r02k_bench_config2_$tag.json 2> gpurun_out/r02k_bench_config2_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/r02k_bench_config2_$tag.json')); print('config2 $tag', round(d['value']/1e6,2), 'M ratings/s', round(d['ms_per_step'],2), 'ms/epoch')"
grep daisy_mf_fit gpurun_out/r02k_bench_config2_$tag.err | tail -2; }
r persm1 DAISY_MF_BLOCKS_PER_SM=1
r poll200 DAISY_MF_POLL_NS=200
r poll1000 DAISY_MF_POLL_NS=1000
r persm1_poll200_seq DAISY_MF_BLOCKS_PER_SM=1 DAISY_MF_POLL_NS=200 DAISY_MF_BATCH=0
