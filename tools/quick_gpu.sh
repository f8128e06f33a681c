cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_cli.py -q -m gpu -x 2>&1 | tail -6
python tools/validate_quality.py 60 fp16x3 2>&1 | tail -12
