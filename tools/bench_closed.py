"""CLOSED insert-or-improve (probe / min / resolve / fix-up launches of dcb_closed_insert) against SURVEY 8(d)'s 44 B/child (8 B hash in
+ 16 B slot probe + 16 B slot write + 4 B keep/slot out), CUDA events, (a) at streaming size for table sizes from L2-resident to
2 GB and (b) at the A* loop's size (240k candidates, half of them duplicates of earlier batches) against the table size."""
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
slot = torch.empty(int(lib.dcb_closed_scratch_bytes(m)), dtype=torch.uint8, device="cuda"); keep = torch.empty(m, dtype=torch.uint8, device="cuda")
counter = torch.zeros(1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
print("# dcb_closed_insert (probe + min + resolve + fix-up launches), %d cube3 children per call, fresh table each call; peak = %.1f GB/s (%s)" % (m, peak, src))
print("%12s %10s %10s %14s %8s %10s" % ("table slots", "table MB", "us", "GB/s (44 B)", "frac", "kept"))
for logcap in (25, 26, 27):          # load 0.37 / 0.19 / 0.09 after the call
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

# (b) the A* loop's shape: 240k candidates per call into a table that already holds `fill` states; half of each call's candidates
# were seen before (the CLOSED hit rate of a cube3 search is 15-60 %)
mb = 240000
print("# in-loop shape: %d candidates per call, ~50%% already in the table" % mb)
print("%12s %10s %10s %10s %14s %8s" % ("table slots", "table MB", "entries", "us", "GB/s (44 B)", "frac"))
for logcap in (21, 23, 25, 27):
    cap = 1 << logcap
    table = torch.empty(cap * 2, dtype=torch.int64, device="cuda")
    _lib.check(lib.dcb_closed_clear(p(table), cap, st)); counter.zero_()
    fill = min(cap // 4, m - 6 * mb)
    for i0 in range(0, fill, 1 << 20):                                 # pre-fill
        n = min(1 << 20, fill - i0)
        _lib.check(lib.dcb_closed_insert(ENV, p(table), cap, p(arena), hs.data_ptr() + 8 * i0, gd.data_ptr() + 4 * i0, None, i0, n, p(slot), p(keep), p(counter), st))
    ts = []
    for it in range(8):
        i0 = fill - mb // 2 + it * (mb // 2)                           # half old, half new
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.dcb_closed_insert(ENV, p(table), cap, p(arena), hs.data_ptr() + 8 * i0, gd.data_ptr() + 4 * i0, None, i0, mb, p(slot), p(keep), p(counter), st))
        b.record(); torch.cuda.synchronize()
        if it >= 2: ts.append(a.elapsed_time(b))
    t = float(np.median(ts)) * 1e-3
    gbs = 44.0 * mb / t / 1e9
    print("%12d %10d %10d %10.1f %14.1f %8.3f" % (cap, cap * 16 >> 20, int(counter.item()), t * 1e6, gbs, gbs / peak))
    del table
