# GEMM A/B on the GPU box: kernel-level parity tests, then network throughput with the CTA-pair (default) and single-CTA kernels.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nnet.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/bench_nnet.py tc 2>&1 | tee gpurun_out/bench_nnet_pair.txt
DCB_GEMM_PAIR=0 timeout 300 python tools/bench_nnet.py "tc fp16x3" 2>&1 | tee gpurun_out/bench_nnet_single.txt
