cd $GRAFT_REPO_ROOT
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1200 python bench.py 2> gpurun_out/bench_$TAG.err > gpurun_out/bench_$TAG.json
tail -5 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.3g e2e %.3g full %.3g  gemm frac %.3f share %.3f" % (d["value"], d["e2e"]["value"], d["full_search"]["value"], d["roofline"]["frac"], d["roofline"]["share_of_timed_region"]))
print(json.dumps(d["other_workloads"], indent=1)[:3000])
print(json.dumps(d["multi_instance"], indent=1))
print(d["cpu_baseline"])
PY
