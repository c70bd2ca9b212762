# Round 2: launch lists (ncu, serialised) of the SGNS and batch-norm BPR-FM steps -- which kernel carries the step.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02u_launches_sgns.csv python bench.py --workload sgns --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02u_ncu_sgns.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02u_launches_bprfm_bn.csv python bench.py --workload bprfm_bn --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02u_ncu_bprfm_bn.log 2>&1
tail -2 gpurun_out/r02u_ncu_sgns.log | cut -c1-300
