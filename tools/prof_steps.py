"""Per-iteration wall time of the search loop on one start state (sync step mode), with CLOSED growth events marked."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from deepcubea_b200.search.bwas_gpu import BWASGpu
wl = sys.argv[1] if len(sys.argv) > 1 else "cube3"
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 60
W = bench.WORKLOADS[wl]
dev = torch.device("cuda")
heur, src = bench.build_heuristic(wl, dev, "fp16x3")
eng = BWASGpu(W["env"], heur, W["weight"], bench.BATCH, max_nodes=1 << 27, device=dev)
states, _ = bench.workload_states(wl, 8)
first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for rep in range(first, first + 2):
    eng.reset(states[rep]); torch.cuda.synchronize()
    rows = []
    for it in range(n_it):
        cap0 = eng.closed_cap
        t0 = time.perf_counter(); eng.step(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ent = int(eng.plan.closed_entries)
        rows.append((it, dt * 1e3, eng.last_popped, eng.last_kept, cap0, eng.closed_cap, ent))
        if eng.done: break
    print("rep", rep)
    prev = 1
    for r in rows:
        new = r[6] - prev; prev = r[6]
        print("  it %3d  %7.2f ms  popped %6d kept %7d (new states %7d, re-opened with smaller g %6d = %4.1f%%)  closed_cap 2^%d%s" % (
            r[0], r[1], r[2], r[3], new, r[3] - new, 100.0 * (r[3] - new) / max(r[3], 1), int(np.log2(r[4])), (" -> 2^%d" % int(np.log2(r[5]))) if r[5] != r[4] else ""))
