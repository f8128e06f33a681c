"""Workload for `ncu -k regex:expand_kernel`: the gather kernel at streaming size (contiguous parents) and at
the A* loop's size (20000 indexed parents)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcubea_b200 import _lib, ops

env = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = 1 << 21
S, A = ops.env_shape(env)
lib = _lib.load()
goal = torch.zeros(S, dtype=torch.uint8)
_lib.check(lib.dcb_env_goal_state(env, goal.data_ptr()))
par = goal.cuda().repeat(n, 1)
g = torch.Generator(device="cuda"); g.manual_seed(0)
for a in torch.randint(0, A, (20,), generator=g, device="cuda").tolist():
    par = ops.next_state(env, par, a)
ch = torch.empty((n, A, S), dtype=torch.uint8, device="cuda")
for _ in range(4):
    ops.expand(env, par, out=ch)
torch.cuda.synchronize()
# indexed variant at the A* batch size
ids = torch.randint(0, n, (20000,), device="cuda", dtype=torch.int32)
sv = torch.empty(20000 * A, dtype=torch.uint8, device="cuda"); hs = torch.empty(20000 * A, dtype=torch.int64, device="cuda")
for _ in range(4):
    _lib.check(lib.dcb_expand_indexed(env, par.data_ptr(), ids.data_ptr(), 20000, ch.data_ptr(), sv.data_ptr(), hs.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("done")
