set -x
python __graft_entry__.py --smoke 2>&1 | tail -3
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_err.log | tee gpurun_out/bench_first.json
tail -5 gpurun_out/bench_err.log
python bench.py --impl reference --steps 4 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref_first.json
