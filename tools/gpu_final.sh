# everything the driver runs at round end, in its order: GPU tests, smoke, reference arm, our arm
cd $GRAFT_REPO_ROOT
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_$TAG.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 2>/dev/null | cut -c1-400
timeout 1200 python bench.py 2> gpurun_out/bench_$TAG.err > gpurun_out/bench_$TAG.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.4g e2e %.4g full %.4g  gemm frac %.3f share %.3f  gather %.3f  cpu %.4g" % (d["value"], d["e2e"]["value"], d["full_search"]["value"], d["roofline"]["frac"], d["roofline"]["share_of_timed_region"], d["roofline_gather"]["frac"], d["cpu_baseline"]["value"]))
print({k: (v.get("value"), v.get("roofline",{}).get("frac"), v.get("roofline_gather",{}).get("frac")) for k,v in d["other_workloads"].items()})
print(d["multi_instance"]["speedup"], d["multi_instance"]["one_engine"]["nodes_per_sec"])
PY
