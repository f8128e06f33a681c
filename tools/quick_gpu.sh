set -x
cd $GRAFT_REPO_ROOT
DCB_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 4 --no_cpu_baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:expand_kernel -c 8 -f -o gpurun_out/prof_expand_r01 python tools/prof_expand.py > gpurun_out/prof_expand.log 2>&1
tail -3 gpurun_out/prof_expand.log
DCB_CUDA_PROFILER=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'closed_|open_|child_meta|compact|gather_nnet|cost_kernel' -c 60 -f -o gpurun_out/prof_bwas_r01 python bench.py --steps 2 --warmup 6 --no_cpu_baseline > gpurun_out/prof_bwas.log 2>&1
tail -3 gpurun_out/prof_bwas.log
ls -la gpurun_out
