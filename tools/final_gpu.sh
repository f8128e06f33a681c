# End-of-round GPU pass: smoke, all GPU tests, bench (ours + reference arm), per-env gather roofline.  Outputs -> gpurun_out/
cd $GRAFT_REPO_ROOT
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py 2>gpurun_out/bench_final.err | tee gpurun_out/bench_final.json | cut -c1-200
python bench.py --impl reference --steps 8 --warmup 3 2>/dev/null | tee gpurun_out/bench_final_reference.json | cut -c1-160
python tools/bench_expand_envs.py 2>&1 | tee gpurun_out/expand_envs_r01.txt
