"""Where does the instrumented bench window lose time?  Same iterations under different instrumentation."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from deepcubea_b200.search.bwas_gpu import BWASGpu
W = bench.WORKLOADS["cube3"]
dev = torch.device("cuda")
heur, src = bench.build_heuristic("cube3", dev, "fp16x3")
eng = BWASGpu(W["env"], heur, W["weight"], bench.BATCH, max_nodes=1 << 27, device=dev)
states, _ = bench.workload_states("cube3", 24)
def run(n_it, first):
    i = first; fresh = True; nodes = 0
    for _ in range(n_it):
        if fresh:
            eng.reset(states[i % len(states)]); fresh = False
        b = eng.nodes_expanded; eng.step(); nodes += eng.nodes_expanded - b
        if eng.done:
            i += 1; fresh = True
    return nodes
def timed(label, n_it, first, gemm=False, expand=False, sampler=False):
    heur.gemm_events = [] if gemm else None
    eng.expand_events = [] if expand else None
    s = None
    if sampler:
        s = bench.ClockSampler(0); s.start()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); nodes = run(n_it, first); e1.record(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if s: s.stop()
    print("%-28s %4d its  wall %7.1f ms  events %7.1f ms  nodes %9d  growths %d" % (label, n_it, dt * 1e3, e0.elapsed_time(e1), nodes, eng.closed_growths), flush=True)
    heur.gemm_events = None; eng.expand_events = None
timed("cold", 47, 0)
timed("plain", 47, 0)
timed("plain again", 47, 0)
timed("gemm events", 47, 0, gemm=True)
timed("expand events", 47, 0, expand=True)
timed("sampler", 47, 0, sampler=True)
timed("all", 47, 0, gemm=True, expand=True, sampler=True)
timed("plain", 47, 0)
