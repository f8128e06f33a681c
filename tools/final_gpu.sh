# End-of-round GPU pass: smoke, all GPU tests, bench (ours + reference arm), ncu launch list of the timed region.  Outputs -> gpurun_out/
cd $GRAFT_REPO_ROOT
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py 2>gpurun_out/bench_final.err | tee gpurun_out/bench_final.json | cut -c1-200
python bench.py --impl reference --steps 8 --warmup 3 2>/dev/null | tee gpurun_out/bench_final_reference.json | cut -c1-160
DCB_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 6 --no_cpu_baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_final.err
