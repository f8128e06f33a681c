cd $GRAFT_REPO_ROOT
for p in 0 2; do echo "== DCB_GEMM_PAIR=$p"; DCB_GEMM_PAIR=$p timeout 120 python tools/bench_nnet.py 2>&1 | tail -2; done
DCB_GEMM_PAIR=2 timeout 300 python -m pytest tests/test_gpu_nnet.py -x -q -m gpu 2>&1 | tail -2
for p in 2 0; do echo "== bench DCB_GEMM_PAIR=$p"; DCB_GEMM_PAIR=$p timeout 200 python bench.py 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline'])"; done
