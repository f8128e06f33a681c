cd $GRAFT_REPO_ROOT
echo "== cube4 default"; timeout 200 python -m pytest tests/test_gpu_env_step.py tests/test_gpu_bwas.py -x -q -m gpu -k cube4 2>&1 | tail -3
echo "== cube4 padded staging"; DCB_CUBE4_PAD=1 timeout 200 python -m pytest tests/test_gpu_env_step.py tests/test_gpu_bwas.py -x -q -m gpu -k cube4 2>&1 | tail -3
echo "== timing"; DCB_BENCH_ENVS=cube4,cube3 timeout 100 python tools/bench_expand_envs.py 2>&1 | tail -3
DCB_CUBE4_PAD=1 DCB_BENCH_ENVS=cube4 timeout 100 python tools/bench_expand_envs.py 2>&1 | tail -1
echo "== regression"; timeout 400 python -m pytest tests/test_gpu_env_step.py tests/test_gpu_nnet.py tests/test_gpu_closed_open.py tests/test_gpu_bwas.py -x -q -m gpu 2>&1 | tail -3
