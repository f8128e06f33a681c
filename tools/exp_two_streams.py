"""Experiment: can the small kernels of one search (pop / expand / CLOSED / push) hide under the GEMMs of ANOTHER search on the same GPU?
Two engines on two CUDA streams, iterations enqueued alternately (each engine one iteration ahead of its records), against the same
two searches run one after the other.  Whole-job view: configs[1] solves many independent instances per GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from deepcubea_b200.search.bwas_gpu import BWASGpu
W = bench.WORKLOADS["cube3"]
dev = torch.device("cuda")
n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 60
states, _ = bench.workload_states("cube3", 8)
pick = [2, 3]                                  # two long searches (see tools/prof_steps.py)
engines, streams, heurs = [], [], []
for k in range(2):
    heur, _ = bench.build_heuristic("cube3", dev, "fp16x3")
    engines.append(BWASGpu(W["env"], heur, W["weight"], bench.BATCH, max_nodes=1 << 26, device=dev))
    streams.append(torch.cuda.Stream(device=dev))

def run_sequential():
    nodes = 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(2):
        eng = engines[k]
        with torch.cuda.stream(streams[k]):
            eng.set_budget(None); eng.reset(states[pick[k]])
            g = eng.pipelined_steps()
            for _ in range(n_it):
                next(g)
                if eng.done: break
            nodes += eng.nodes_expanded
        torch.cuda.synchronize()
    return nodes, time.perf_counter() - t0

def run_interleaved():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    gens = []
    for k in range(2):
        with torch.cuda.stream(streams[k]):
            engines[k].set_budget(None); engines[k].reset(states[pick[k]])
            gens.append(engines[k].pipelined_steps())
    live = [True, True]
    for _ in range(n_it):
        for k in range(2):
            if live[k]:
                with torch.cuda.stream(streams[k]):
                    next(gens[k])
                if engines[k].done: live[k] = False
    torch.cuda.synchronize()
    return sum(e.nodes_expanded for e in engines), time.perf_counter() - t0

for name, fn in (("warm", run_sequential), ("sequential", run_sequential), ("interleaved (2 streams)", run_interleaved), ("sequential", run_sequential),
                 ("interleaved (2 streams)", run_interleaved)):
    nodes, sec = fn()
    print("%-26s %10d nodes  %8.1f ms  %6.2f M nodes/s" % (name, nodes, sec * 1e3, nodes / sec / 1e6), flush=True)
