cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 30 --warmup 6 --no_cpu_baseline 2>gpurun_out/bench_2gpu.err | tee gpurun_out/bench_2gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('N=2 value %.4g ms/step %.2f e2e %.4g n_gpus %d scaling %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'], d['scaling']))"
tail -3 gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-160
python bench.py --steps 30 --warmup 6 --no_cpu_baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('N=1 value %.4g ms/step %.2f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
