cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# warm caches as in the running loop: no flush between kernels, no clock control
timeout 900 ncu --set full --clock-control none --cache-control none -k regex:"closed_|search_push|search_plan|onehot_gather|expand_kernel" -s 200 -c 16 -o gpurun_out/loop_warm_r02 -f python tools/prof_steps.py cube3 45 > gpurun_out/ncu_loop.log 2>&1
ls -la gpurun_out/*.ncu-rep
