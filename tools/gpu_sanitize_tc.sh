cd $GRAFT_REPO_ROOT
for pair in 1 0; do
for tool in synccheck racecheck; do
  echo "=== DCB_GEMM_PAIR=$pair compute-sanitizer --tool $tool python tools/sanitize.py tc"
  DCB_GEMM_PAIR=$pair timeout 900 compute-sanitizer --tool $tool python tools/sanitize.py tc 2>&1 | grep -v "Host Frame\|^=========         in \|Saved host" | grep -E "error detected|Race reported|    at |    and |by thread|SUMMARY|tc net" | head -24
done; done
timeout 600 python -m pytest tests/test_gpu_nnet.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_nnet.py "tc fp16x3" 2>&1 | tail -3
