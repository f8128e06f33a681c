"""Time the cube3 gather kernel (2^21 parents, CUDA events) for every DCB_EXPAND_CFG shape, one subprocess each."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
from deepcubea_b200 import ops
n = 1 << 21
par = torch.arange(54, dtype=torch.uint8, device="cuda").repeat(n, 1)
g = torch.Generator(device="cuda"); g.manual_seed(0)
for a in torch.randint(0, 12, (12,), generator=g, device="cuda").tolist(): par = ops.next_state(0, par, a)
ch = torch.empty((n, 12, 54), dtype=torch.uint8, device="cuda")
sv = None
ts = []
for it in range(12):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); c, sv, hs = ops.expand(0, par, out=ch); b.record(); torch.cuda.synchronize()
    if it >= 4: ts.append(a.elapsed_time(b))
t = float(np.median(ts)) * 1e-3
print(json.dumps({"cfg": os.environ.get("DCB_EXPAND_CFG", "default"), "us": t * 1e6, "GBs": 67.5 * n * 12 / t / 1e9, "chk": int(c[::4097].sum().item()), "hchk": int(hs[::4097].sum().item())}))
''' % ROOT
for cfg in ["4x1", "5x1", "1x2", "2x2", "5x2"]:
    env = dict(os.environ, DCB_EXPAND_CFG=cfg)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(out.stdout.strip().split("\n")[-1] if out.returncode == 0 else ("FAILED " + cfg + out.stderr[-300:]))
