cd $GRAFT_REPO_ROOT
timeout 180 python -m pytest tests/test_gpu_nnet.py -q -m gpu -x 2>&1 | tail -3
timeout 120 python tools/bench_nnet.py 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_closed_open.py -q -m gpu -x 2>&1 | tail -12
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
