# GPU parity tests only.  Usage on the box: bash tools/gpu_tests.sh [pytest args]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu "$@" 2>&1 | tail -40 | tee gpurun_out/pytest_last.txt
