"""Solution quality on the reference's own test set, against the reference's shipped results and the optimal lengths.
    python tools/validate_quality.py [n_states] [precision]   -> prints a table, writes profiles/quality_r01_cube3.txt
Config = the reference's published cube3 run (train.sh:9): weight 0.6, batch_size 10000."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from deepcubea_b200.nnet.folded import DeviceHeuristic, FoldedResnet
from deepcubea_b200.nnet.tc_resnet import TcResnet
from deepcubea_b200.search.bwas_gpu import BWASGpu
from deepcubea_b200.utils.env_utils import get_environment
from deepcubea_b200.utils.nnet_utils import load_nnet
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
env = get_environment("cube3")
model = load_nnet(os.path.join(ROOT, "assets/saved_models/cube3/current/model_state_dict.pt"), env.get_nnet_model(), device=torch.device("cpu"))
dev = torch.device("cuda")
heur = TcResnet(model, dev, prec) if prec in ("fp16x3", "fp16") else DeviceHeuristic(FoldedResnet(model, prec).to(dev))
g = np.load(os.path.join(ROOT, "tests/golden/optimal_cube3.npz")); ref = np.load(os.path.join(ROOT, "tests/golden/paths_cube3.npz"))
opt = np.diff(g["offsets"]); ref_len = np.diff(ref["offsets"]) - 1
eng = BWASGpu("cube3", heur, 0.6, 10000, max_nodes=1 << 27)
lens, nodes, secs = [], [], []
for i in range(n):
    r = eng.solve(g["states"][i])
    cur = torch.from_numpy(g["states"][i][None]).cuda()
    from deepcubea_b200 import ops
    for mv in r.moves: cur = ops.next_state(0, cur, mv)
    assert bool(ops.is_solved(0, cur)[0]), i
    lens.append(len(r.moves)); nodes.append(r.nodes_generated); secs.append(r.solve_time)
lens, nodes, secs = np.array(lens), np.array(nodes), np.array(secs)
lines = ["# cube3, weight 0.6, batch_size 10000 (reference train.sh:9), first %d states of data/cube3/test, heuristic %s" % (n, prec),
         "# ours vs the reference's shipped results/cube3/results.pkl on the SAME states, and vs the optimal lengths of data/cube3/test",
         "mean length          ours %.3f   reference %.3f   optimal %.3f" % (lens.mean(), ref_len[:n].mean(), opt[:n].mean()),
         "%% optimal            ours %.1f   reference %.1f" % (100 * np.mean(lens == opt[:n]), 100 * np.mean(ref_len[:n] == opt[:n])),
         "max gap to optimal   ours %d   reference %d" % ((lens - opt[:n]).max(), (ref_len[:n] - opt[:n]).max()),
         "ours == reference length: %.1f%%; ours shorter %d, longer %d" % (100 * np.mean(lens == ref_len[:n]), int((lens < ref_len[:n]).sum()), int((lens > ref_len[:n]).sum())),
         "mean nodes generated ours %.3g   reference %.3g" % (nodes.mean(), ref["num_nodes_generated"][:n].mean()),
         "mean solve time (s)  ours %.3f   reference %.2f (unstated hardware)" % (secs.mean(), ref["times"][:n].mean()),
         "nodes/s (sum/sum)    ours %.4g   reference %.4g" % (nodes.sum() / secs.sum(), ref["num_nodes_generated"][:n].sum() / ref["times"][:n].sum()),
         "all %d solutions valid (replayed through dcb_next_state / dcb_is_solved)" % n]
print("\n".join(lines))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "quality_r01_cube3_%s.txt" % prec), "w").write("\n".join(lines) + "\n")
