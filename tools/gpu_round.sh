# Full GPU check: parity tests, bench (both arms), launch list.  Usage on the box: bash tools/gpu_round.sh [tag]
cd $GRAFT_REPO_ROOT
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_$TAG.txt
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 2> gpurun_out/bench_ref_$TAG.err | tee gpurun_out/bench_ref_$TAG.json
timeout 900 python bench.py 2> gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json
tail -5 gpurun_out/bench_$TAG.err
timeout 300 python tools/bench_nnet.py tc 2>&1 | tee gpurun_out/bench_nnet_$TAG.txt
