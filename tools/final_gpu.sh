# End-of-round GPU pass: tests, bench (ours + reference arm), ncu launch list + full captures.  Outputs -> gpurun_out/
cd $GRAFT_REPO_ROOT
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py 2>gpurun_out/bench_final.err | tee gpurun_out/bench_final.json | cut -c1-300
python bench.py --impl reference --steps 8 --warmup 3 2>/dev/null | tee gpurun_out/bench_final_reference.json | cut -c1-200
python bench.py --nnet_precision fp32 --steps 12 --warmup 4 --no_cpu_baseline 2>/dev/null | tee gpurun_out/bench_final_fp32.json | cut -c1-200
DCB_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 6 --no_cpu_baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resnet_gemm -s 7 -c 7 -f -o gpurun_out/prof_gemm_final python tools/prof_gemm.py > gpurun_out/prof_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:expand_kernel -c 8 -f -o gpurun_out/prof_expand_final python tools/prof_expand.py > gpurun_out/prof_expand.log 2>&1
ls gpurun_out | head -30
