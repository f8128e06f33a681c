"""CLOSED insert + resolve at streaming size against SURVEY 8(d)'s 44 B/child (8 B hash in + 16 B slot probe + 16 B slot write
+ 4 B keep/slot out), CUDA events, for table sizes from L2-resident to 1 GB.  Written at the end of r01 after the GPU budget was
spent: not yet run on a B200 (the in-loop figures in DESIGN.md come from profiles/bwas_kernels_r01_ncu.txt)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from deepcubea_b200 import _lib, ops
lib = _lib.load(); p = _lib.ptr
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]; src = "measured"
except Exception:
    peak, src = 6650.0, "fallback"
ENV, S, A = 0, 54, 12
n_par = 1 << 20
par = torch.arange(S, dtype=torch.uint8, device="cuda").repeat(n_par, 1)
g = torch.Generator(device="cuda"); g.manual_seed(1)
idx = torch.arange(n_par, device="cuda")
for step in range(14):                                   # diversify: every parent takes its own move sequence
    mv = torch.randint(0, A, (n_par,), generator=g, device="cuda")
    for a in range(A):
        sel = mv == a
        par[sel] = ops.next_state(ENV, par[sel].contiguous(), a)
ch, _, hs = ops.expand(ENV, par)
m = n_par * A
arena = torch.empty(m * S + 64, dtype=torch.uint8, device="cuda"); arena[:m * S] = ch.reshape(-1)
hs = hs.reshape(-1).contiguous()
gd = torch.ones(m, dtype=torch.int32, device="cuda")
slot = torch.empty(m, dtype=torch.int32, device="cuda"); keep = torch.empty(m, dtype=torch.uint8, device="cuda")
counter = torch.zeros(1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
print("# dcb_closed_insert (insert + resolve launches), %d cube3 children per call, fresh table each call; peak = %.1f GB/s (%s)" % (m, peak, src))
print("%12s %10s %10s %14s %8s %10s" % ("table slots", "table MB", "us", "GB/s (44 B)", "frac", "kept"))
for logcap in (25, 26, 27):
    cap = 1 << logcap
    table = torch.empty(cap * 2, dtype=torch.int64, device="cuda")
    ts = []
    for it in range(6):
        _lib.check(lib.dcb_closed_clear(p(table), cap, st)); counter.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.dcb_closed_insert(ENV, p(table), cap, p(arena), p(hs), p(gd), None, 0, m, p(slot), p(keep), p(counter), st))
        b.record(); torch.cuda.synchronize()
        if it >= 2: ts.append(a.elapsed_time(b))
    t = float(np.median(ts)) * 1e-3
    gbs = 44.0 * m / t / 1e9
    print("%12d %10d %10.1f %14.1f %8.3f %10d" % (cap, cap * 16 >> 20, t * 1e6, gbs, gbs / peak, int(keep.sum())))
    del table
