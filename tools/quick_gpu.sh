cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_nnet.py -q -m gpu -x 2>&1 | tail -2
timeout 120 python tools/bench_nnet.py 2>&1 | tail -2
DCB_GEMM_PAIR=1 timeout 120 python tools/bench_nnet.py 2>&1 | tail -2
DCB_GEMM_PAIR=2 timeout 120 python tools/bench_nnet.py 2>&1 | tail -2
