cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -m gpu -k two_ranks > gpurun_out/cli2.log 2>&1
tail -3 gpurun_out/cli2.log
grep -n "Error\|error\|Traceback" gpurun_out/cli2.log | head -10
