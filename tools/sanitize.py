"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every hand-written kernel once, sizes chosen so
edge paths run (partial tiles, odd counts, table probes, pop with ties, GEMM with partial M tile, K chunks, residual)."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepcubea_b200 import _lib, ops
from deepcubea_b200.nnet.tc_resnet import TcResnet
from deepcubea_b200.search.bwas_gpu import BWASGpu
from deepcubea_b200.utils.pytorch_models import ResnetModel
lib = _lib.load()
g = torch.Generator(device="cuda"); g.manual_seed(0)
for env, name in enumerate(["cube3", "puzzle15", "puzzle24", "puzzle35", "puzzle48", "lightsout7", "cube4"]):
    S, A = ops.env_shape(env)
    goal = torch.zeros(S, dtype=torch.uint8); _lib.check(lib.dcb_env_goal_state(env, goal.data_ptr()))
    st = goal.cuda().repeat(333, 1)
    for a in torch.randint(0, A, (9,), generator=g, device="cuda").tolist():
        st = ops.next_state(env, st, a)
    ch, sv, hs = ops.expand(env, st)
    ops.is_solved(env, st); ops.hash_states(env, st); ops.nnet_input(env, st)
    goal_in = ops.nnet_input(env, goal.cuda()[None])[0]
    eng = BWASGpu(name, lambda x, gi=goal_in: (x != gi[None]).sum(dim=1).float() / 8.0, 0.8, 37, max_nodes=1 << 16)
    r = eng.solve(st[0].cpu().numpy(), max_iters=12)
    print(name, "expand", tuple(ch.shape), "bwas iterations", r.iterations, "nodes", r.nodes_generated)
torch.manual_seed(0)
tc = TcResnet(ResnetModel(16, 16, 300, 200, 1, 1, True).eval(), torch.device("cuda"), "fp16x3", chunk=1000)
x = torch.randint(0, 16, (333, 16), device="cuda", dtype=torch.uint8)
print("tc net", float(tc(x).sum()))
torch.cuda.synchronize()
print("sanitize workload done")
