# End-of-round GPU pass: tests, bench (ours + reference arm), ncu launch list + full captures, puzzle48 validation.
cd $GRAFT_REPO_ROOT
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py 2>gpurun_out/bench_final.err | tee gpurun_out/bench_final.json | cut -c1-240
python bench.py --impl reference --steps 8 --warmup 3 2>/dev/null | tee gpurun_out/bench_final_reference.json | cut -c1-160
DCB_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 6 --no_cpu_baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resnet_gemm -s 6 -c 6 -f -o gpurun_out/prof_gemm_final python tools/prof_gemm.py > gpurun_out/prof_gemm.log 2>&1
if [ -f assets/saved_models/puzzle48/current/model_state_dict.pt ]; then timeout 600 python tools/validate_quality.py puzzle48 20 fp16x3 2>&1 | tail -9; fi
ls gpurun_out | head -30
