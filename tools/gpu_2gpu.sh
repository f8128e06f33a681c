cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no_cpu_baseline 2> gpurun_out/bench_${N}gpu.err | grep "^{" > gpurun_out/bench_r02_${N}gpu.json
tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02_${N}gpu.json"))
print("N=$N value %.4g e2e %.4g full %.4g ms/step %.2f share %.3f" % (d["value"], d["e2e"]["value"], d["full_search"]["value"], d["ms_per_step"], d["roofline"]["share_of_timed_region"]))
for r in d["config"]["per_rank"]: print(r)
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --full --num_states 30 2>> gpurun_out/bench_${N}gpu.err | grep "^{" > gpurun_out/bench_full_${N}gpu_small.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_full_${N}gpu_small.json"))
c=d.pop("config")
print({k:d[k] for k in ("value","solved","unsolved","nodes_generated","wall_s","balance")}, c["per_rank"], c.get("lanes_per_gpu"))
PY
