"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every hand-written kernel once, sizes chosen so
edge paths run (partial tiles, odd counts, table probes, pop with ties, GEMM with partial M tile, K chunks, residual)."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepcubea_b200 import _lib, ops
from deepcubea_b200.nnet.tc_resnet import TcResnet
from deepcubea_b200.search.engine import BWASGpu, SearchEngine
from deepcubea_b200.utils.pytorch_models import ResnetModel
lib = _lib.load()
g = torch.Generator(device="cuda"); g.manual_seed(0)
only_tc = len(sys.argv) > 1 and sys.argv[1] == "tc"
for env, name in enumerate([] if only_tc else ["cube3", "puzzle15", "puzzle24", "puzzle35", "puzzle48", "lightsout7", "cube4"]):
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
    # the sync-free pipelined loop (device-side counts end to end) and a many-instance engine with the Python semantics
    easy = goal.cuda()[None]
    for a in torch.randint(0, A, (3,), generator=g, device="cuda").tolist():
        easy = ops.next_state(env, easy, a)
    eng2 = BWASGpu(name, lambda x, gi=goal_in: (x != gi[None]).sum(dim=1).float() / 8.0, 0.8, 37, max_nodes=1 << 20, sync_free=True)
    r2 = eng2.solve(easy[0].cpu().numpy())
    multi = SearchEngine(name, lambda x, gi=goal_in: (x != gi[None]).sum(dim=1).float() / 8.0, [0.5] * 7, 5, n_inst=7, max_nodes=1 << 17,
                         semantics="python", sync_free=True)
    multi.raise_on_error = False
    multi.reset(st[2:9].cpu().numpy())
    for _ in range(6):
        multi.step_all()
    print(name, "expand", tuple(ch.shape), "bwas iterations", r.iterations, "nodes", r.nodes_generated, "| pipelined", r2.iterations, r2.nodes_generated,
          "| 7 instances, 6 steps:", sum(int(x.nodes_generated) for x in multi.inst))
if len(sys.argv) > 1 and sys.argv[1] == "search":            # search-loop kernels only
    torch.cuda.synchronize(); print("sanitize workload done (search kernels only)"); sys.exit(0)
torch.manual_seed(0)
tc = TcResnet(ResnetModel(16, 16, 300, 200, 1, 1, True).eval(), torch.device("cuda"), "fp16x3", chunk=1000)
x = torch.randint(0, 16, (333, 16), device="cuda", dtype=torch.uint8)
print("tc net", float(tc(x).sum()))
ids = torch.arange(333, dtype=torch.int32, device="cuda"); cnt = torch.tensor([200], dtype=torch.int32, device="cuda")
arena = x.reshape(-1).contiguous()
kind, *res = tc.eval_nodes_dev(1, arena, ids, cnt.data_ptr(), 333)          # device-side row count through every layer
print("tc net, device row count:", kind, float(res[0][:200].sum()))
torch.cuda.synchronize()
print("sanitize workload done")
