cd $GRAFT_REPO_ROOT
N=${1:-8}; NS=${2:-100}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no_cpu_baseline 2> gpurun_out/bench_${N}gpu.err > gpurun_out/bench_r02_${N}gpu.json
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_r02_${N}gpu.json") if l.startswith("{")][0]
print("N=$N value %.4g e2e %.4g full %.4g ms/step %.2f clocks %s" % (d["value"], d["e2e"]["value"], d["full_search"]["value"], d["ms_per_step"], d["clocks"]))
for r in d["config"]["per_rank"]: print(r)
PY
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --full --num_states $NS 2> gpurun_out/bench_full_${N}gpu.err > gpurun_out/bench_full_r02_${N}gpu.json
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_full_r02_${N}gpu.json") if l.startswith("{")][0]
c=d.pop("config")
print({k:d[k] for k in ("value","solved","unsolved","nodes_generated","wall_s","mean_solution_len","balance")})
for r in c["per_rank"]: print(r)
PY
tail -2 gpurun_out/bench_full_${N}gpu.err
