"""Solution quality on the reference's own test sets, against the reference's shipped results and the optimal lengths.
    python tools/validate_quality.py <env> [n_states] [precision]  -> prints a table, writes gpurun_out/quality_r02_<env>_<prec>.txt
Configs = the reference's published runs (train.sh:9 cube3: weight 0.6, batch 10000; train.sh:21 puzzle15: 0.8 / 20000;
train.sh:57 puzzle48: 0.6 / 20000; train.sh:68 lightsout7: 0.2 / 1000).  Needs assets/saved_models/<env>/current/model_state_dict.pt (tools/fetch_assets.py <env>)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from deepcubea_b200 import _lib, ops
from deepcubea_b200.nnet.folded import DeviceHeuristic, FoldedResnet
from deepcubea_b200.nnet.tc_resnet import TcResnet
from deepcubea_b200.search.bwas_gpu import BWASGpu
from deepcubea_b200.utils.env_utils import get_environment
from deepcubea_b200.utils.nnet_utils import load_nnet
name = sys.argv[1] if len(sys.argv) > 1 else "cube3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
prec = sys.argv[3] if len(sys.argv) > 3 else "fp16x3"
weight, batch = {"cube3": (0.6, 10000), "puzzle15": (0.8, 20000), "puzzle48": (0.6, 20000), "lightsout7": (0.2, 1000)}[name]
env = get_environment(name); eid = _lib.ENV_IDS[name]
model = load_nnet(os.path.join(ROOT, "assets/saved_models/%s/current/model_state_dict.pt" % name), env.get_nnet_model(), device=torch.device("cpu"))
dev = torch.device("cuda")
heur = TcResnet(model, dev, prec) if prec in ("fp16x3", "fp16") else DeviceHeuristic(FoldedResnet(model, prec).to(dev))
ref = np.load(os.path.join(ROOT, "tests/golden/paths_%s.npz" % name)); ref_len = np.diff(ref["offsets"]) - 1
starts = ref["states"][ref["offsets"][:-1]]
optf = os.path.join(ROOT, "tests/golden/optimal_%s.npz" % name)
opt = np.diff(np.load(optf)["offsets"]) if os.path.exists(optf) else None
if opt is not None:
    assert np.array_equal(np.load(optf)["states"][:n], starts[:n])
eng = BWASGpu(name, heur, weight, batch, max_nodes=1 << 27)
lens, nodes, secs = [], [], []
for i in range(n):
    r = eng.solve(starts[i])
    cur = torch.from_numpy(starts[i][None]).cuda()
    for mv in r.moves: cur = ops.next_state(eid, cur, mv)
    assert bool(ops.is_solved(eid, cur)[0]), i
    lens.append(len(r.moves)); nodes.append(r.nodes_generated); secs.append(r.solve_time)
lens, nodes, secs = np.array(lens), np.array(nodes), np.array(secs)
rn = ref["num_nodes_generated"][:n]
lines = ["# %s, weight %s, batch_size %d (the reference's published config), first %d states of data/%s/test, heuristic %s" % (name, weight, batch, n, name, prec),
         "# ours vs the reference's shipped results/%s/results.pkl on the SAME states" % name + (", and vs the optimal lengths of data/%s/test" % name if opt is not None else ""),
         "mean length          ours %.3f   reference %.3f" % (lens.mean(), ref_len[:n].mean()) + ("   optimal %.3f" % opt[:n].mean() if opt is not None else "")]
if opt is not None:
    lines += ["%% optimal            ours %.1f   reference %.1f" % (100 * np.mean(lens == opt[:n]), 100 * np.mean(ref_len[:n] == opt[:n])),
              "max gap to optimal   ours %d   reference %d" % ((lens - opt[:n]).max(), (ref_len[:n] - opt[:n]).max())]
lines += ["ours == reference length: %.1f%%; ours shorter %d, longer %d" % (100 * np.mean(lens == ref_len[:n]), int((lens < ref_len[:n]).sum()), int((lens > ref_len[:n]).sum())),
          "nodes generated: mean ours %.4g   reference %.4g; identical count on %d / %d states; median |ours-ref|/ref %.2e" % (
              nodes.mean(), rn.mean(), int((nodes == rn).sum()), n, float(np.median(np.abs(nodes - rn) / rn))),
          "mean solve time (s)  ours %.3f   reference %.2f (unstated hardware)" % (secs.mean(), ref["times"][:n].mean()),
          "nodes/s (sum/sum)    ours %.4g   reference %.4g   ratio %.1fx" % (nodes.sum() / secs.sum(), rn.sum() / ref["times"][:n].sum(), (nodes.sum() / secs.sum()) / (rn.sum() / ref["times"][:n].sum())),
          "all %d solutions valid (replayed through dcb_next_state / dcb_is_solved)" % n]
print("\n".join(lines))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "quality_r02_%s_%s.txt" % (name, prec)), "w").write("\n".join(lines) + "\n")
