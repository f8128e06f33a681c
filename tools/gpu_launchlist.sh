# launch list (ncu, cold-cache serialised times: shares only) of a short bench window + the bench itself
cd $GRAFT_REPO_ROOT
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python bench.py --no_cpu_baseline 2> gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json
tail -3 gpurun_out/bench_$TAG.err
DCB_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 4 --warmup 3 --no_cpu_baseline --full_states 0 > gpurun_out/launches_$TAG.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_$TAG.csv | tee gpurun_out/launches_${TAG}_summary.txt
