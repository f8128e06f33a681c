cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu -s -x 2>&1 | grep -E "max \|err\||assert|passed|failed|Error" | head -12
python tools/bench_nnet.py 2>&1 | tail -5
for p in fp16x3 fp16; do python bench.py --steps 30 --warmup 8 --no_cpu_baseline --nnet_precision $p 2>gpurun_out/bench_$p.err | tee gpurun_out/bench_$p.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['dtype'][-40:], 'value %.3g'%d['value'], 'ms/step %.2f'%d['ms_per_step'], 'e2e %.3g'%d['e2e']['value'], 'solved', d['config']['solved_in_timed_region'], 'len', d['config']['mean_solution_len'])"; tail -2 gpurun_out/bench_$p.err; done
