cd $GRAFT_REPO_ROOT
for kc in 1024 2048 5120; do echo "K_CHUNK=$kc"; DCB_K_CHUNK=$kc timeout 200 python -m pytest tests/test_gpu_nnet.py -q -m gpu -s -k "network or golden" 2>&1 | grep -E "max \|err\||passed|failed|assert" | head -4; done
timeout 120 python tools/bench_nnet.py 2>&1 | tail -2
DCB_K_CHUNK=5120 timeout 120 python tools/bench_nnet.py 2>&1 | tail -2
