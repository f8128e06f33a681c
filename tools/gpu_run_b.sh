cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_scale.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -15
timeout 900 python bench.py --no_cpu_baseline 2> gpurun_out/bench_r02b.err | tee gpurun_out/bench_r02b.json
tail -3 gpurun_out/bench_r02b.err
timeout 600 python tools/validate_quality.py lightsout7 60 2>&1 | tail -8
timeout 600 python tools/validate_quality.py cube3 30 2>&1 | tail -8
