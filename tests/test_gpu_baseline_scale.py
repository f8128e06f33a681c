"""GPU, BASELINE scale: whole searches at the reference's PUBLISHED configurations (train.sh:9 cube3 weight 0.6 / batch 10000;
:21 puzzle15 0.8 / 20000; :57 puzzle48 0.6 / 20000; :68 lightsout7 0.2 / 1000) with the reference's trained networks on the
hand-written tcgen05 path, on start states of the reference's own test sets -- the number of nodes generated and the solution
length must EQUAL what the reference shipped for that state (results/<env>/results.pkl -> tests/golden/paths_<env>.npz), and the
solution must replay to the goal.  The states with the smallest shipped searches are used so the tier stays short."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFIGS = [("cube3", 0.6, 10000, 3, 1 << 25), ("puzzle15", 0.8, 20000, 3, 1 << 25), ("puzzle48", 0.6, 20000, 1, 1 << 26),
           ("lightsout7", 0.2, 1000, 3, 1 << 24)]


@pytest.mark.parametrize("name,weight,batch,n_states,max_nodes", CONFIGS)
def test_published_config_reproduces_shipped_node_counts(golden_dir, name, weight, batch, n_states, max_nodes):
    wfile = os.path.join(ROOT, "assets", "saved_models", name, "current", "model_state_dict.pt")
    if not os.path.exists(wfile):
        pytest.skip("trained weights not present (tools/fetch_assets.py %s)" % name)
    from deepcubea_b200 import _lib, ops
    from deepcubea_b200.nnet.tc_resnet import TcResnet
    from deepcubea_b200.search.bwas_gpu import BWASGpu
    from deepcubea_b200.utils.env_utils import get_environment
    from deepcubea_b200.utils.nnet_utils import load_nnet
    ref = np.load(golden_dir + "/paths_%s.npz" % name)
    ref_len = np.diff(ref["offsets"]) - 1
    starts = ref["states"][ref["offsets"][:-1]]
    order = np.argsort(ref["num_nodes_generated"], kind="stable")[:n_states]
    env = get_environment(name)
    eid = _lib.ENV_IDS[name]
    dev = torch.device("cuda")
    model = load_nnet(wfile, env.get_nnet_model(), device=torch.device("cpu"))
    heur = TcResnet(model, dev, "fp16x3")
    eng = BWASGpu(name, heur, weight, batch, max_nodes=max_nodes)
    launches0 = heur.gemm_launches
    for i in order:
        r = eng.solve(starts[i])                   # device-driven, pipelined: no host round trip inside an iteration
        assert r.moves is not None
        cur = torch.from_numpy(starts[i][None]).cuda()
        for mv in r.moves:
            cur = ops.next_state(eid, cur, mv)
        assert bool(ops.is_solved(eid, cur)[0]), "state %d: invalid solution" % i
        assert r.nodes_generated == int(ref["num_nodes_generated"][i]), "state %d: nodes generated %d, reference shipped %d" % (
            i, r.nodes_generated, int(ref["num_nodes_generated"][i]))
        assert len(r.moves) == int(ref_len[i]), "state %d" % i
    assert heur.gemm_launches > launches0          # the heuristic ran on the hand-written tcgen05 layers


def test_astar_class_many_instances_with_trained_network(golden_dir):
    """AStar(states, env, heuristic_fn, weights) with 32 cube3 instances and the trained network: every instance is solved with a
    valid path, with ONE heuristic evaluation per step for all instances together (search_methods/astar.py:256-317)."""
    wdir = os.path.join(ROOT, "assets", "saved_models", "cube3", "current")
    if not os.path.exists(os.path.join(wdir, "model_state_dict.pt")):
        pytest.skip("trained weights not present")
    import random
    from deepcubea_b200.search_methods.astar import AStar, get_path
    from deepcubea_b200.utils import nnet_utils
    from deepcubea_b200.utils.env_utils import get_environment
    from deepcubea_b200.utils.search_utils import is_valid_soln
    env = get_environment("cube3")
    np.random.seed(5); random.seed(5)
    states, depths = env.generate_states(32, (4, 12))
    device, _, on_gpu = nnet_utils.get_device()
    fn = nnet_utils.load_heuristic_fn(wdir, device, on_gpu, env.get_nnet_model(), env, clip_zero=True)
    astar = AStar(states, env, fn, [0.6] * 32, max_nodes=1 << 25)
    steps = 0
    while not min(astar.has_found_goal()):
        astar.step(fn, 200)
        steps += 1
        assert steps < 400
    assert astar.heuristic_calls == steps + 1            # the roots, then one call per step
    assert fn.device_fn.gemm_launches > 0
    for i, s in enumerate(states):
        path, soln, cost = get_path(astar.get_goal_node_smallest_path_cost(i))
        assert is_valid_soln(s, soln, env) and cost == len(soln) <= max(depths[i], 1) + 6 and path[0] == s
        assert astar.get_num_nodes_generated(i) > 0
