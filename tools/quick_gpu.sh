cd $GRAFT_REPO_ROOT
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gpu_env_step.py tests/test_gpu_bwas.py -x -q -m gpu -k cube4 2>&1 | tail -2
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize.py 2>&1 | grep -E "ERROR SUMMARY|Invalid|cube4|done" | head -8
