"""Gather-kernel roofline for every environment (algorithmic bytes per child of SURVEY 8d: S/A + S + 1 + 8), CUDA events,
outputs larger than L2.  Prints a table; bench.py reports the cube3 row as roofline_gather."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from deepcubea_b200 import _lib, ops
lib = _lib.load()
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]; src = "measured"
except Exception:
    peak, src = 6650.0, "fallback"
print("# expand + is_solved + hash kernel, one launch, CUDA events (median of 8 after 4 warm-ups); peak = %.1f GB/s (%s copy bandwidth)" % (peak, src))
print("%-11s %10s %6s %4s %12s %10s %12s %8s" % ("env", "parents", "S", "A", "B/child", "us", "GB/s (alg.)", "frac"))
only = os.environ.get("DCB_BENCH_ENVS")          # e.g. "cube4" to time one environment
for env, name in enumerate(["cube3", "puzzle15", "puzzle24", "puzzle35", "puzzle48", "lightsout7", "cube4"]):
    if only and name not in only.split(","):
        continue
    S, A = ops.env_shape(env)
    alg = S / A + S + 1 + 8
    n = int(min(1 << 23, (1.6e9 // (A * S)) // 1024 * 1024))
    goal = torch.zeros(S, dtype=torch.uint8); _lib.check(lib.dcb_env_goal_state(env, goal.data_ptr()))
    par = goal.cuda().repeat(n, 1)
    g = torch.Generator(device="cuda"); g.manual_seed(env)
    for a in torch.randint(0, A, (14,), generator=g, device="cuda").tolist():
        par = ops.next_state(env, par, a)
    ch = torch.empty((n, A, S), dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.expand(env, par, out=ch); b.record(); torch.cuda.synchronize()
        if it >= 4: ts.append(a.elapsed_time(b))
    t = float(np.median(ts)) * 1e-3
    gbs = alg * n * A / t / 1e9
    print("%-11s %10d %6d %4d %12.2f %10.1f %12.1f %8.3f" % (name, n, S, A, alg, t * 1e6, gbs, gbs / peak))
    del par, ch
