# ncu evidence: CLOSED kernels inside the A* loop (DRAM traffic per candidate), GEMM layers at 131072 rows, bench_closed table
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/bench_closed.py 2>&1 | tee gpurun_out/closed_r02.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:closed_ -s 120 -c 8 -o gpurun_out/closed_r02 -f python tools/prof_steps.py cube3 45 > gpurun_out/ncu_closed.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:resnet_gemm_pair -s 20 -c 10 -o gpurun_out/gemm_r02 -f python tools/bench_nnet.py "tc fp16x3" > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:expand_kernel -s 40 -c 3 -o gpurun_out/expand_inloop_r02 -f python tools/prof_steps.py cube3 45 > gpurun_out/ncu_expand.log 2>&1
ls -la gpurun_out/*.ncu-rep
