cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_nnet.py tests/test_gpu_bwas.py tests/test_gpu_closed_open.py -q -m gpu -x 2>&1 | tail -2
timeout 120 python tools/bench_nnet.py 2>&1 | tail -2
python bench.py --steps 30 --warmup 8 --no_cpu_baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value %.4g ms/step %.2f e2e %.4g roof %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], json.dumps({k:d['roofline'][k] for k in ('achieved','peak','frac','share_of_timed_region','avg_us')})))"
