"""GPU: the drop-in surface end to end -- Environment API on host objects, AStar class, astar CLI + results.pkl."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WEIGHTS_DIR = os.path.join(ROOT, "assets", "saved_models", "cube3", "current")


def test_environment_api_matches_oracle():
    import random
    from deepcubea_b200.utils.env_utils import get_environment
    from deepcubea_b200.utils.search_utils import is_valid_soln
    from oracle import oracle_env as O
    for name in ("cube3", "puzzle15", "puzzle48", "lightsout7"):
        env, orc = get_environment(name), O.get_oracle_env(name)
        np.random.seed(3); random.seed(3)
        states, depths = env.generate_states(200, (0, 12))
        np.random.seed(3); random.seed(3)
        ost, odepths = orc.generate_states(200, (0, 12))
        assert np.array_equal(env.pack(states), ost) and list(depths) == list(odepths)     # same RNG call sequence as the reference
        exp, tcs = env.expand(states)
        och, _ = orc.expand(ost)
        assert np.array_equal(np.stack([env.pack(row) for row in exp]), och)
        assert all((tc == 1.0).all() for tc in tcs)
        assert np.array_equal(env.is_solved(states), orc.is_solved(ost))
        assert np.array_equal(env.state_to_nnet_input(states)[0], orc.nnet_input(ost))
        nxt, tc = env.next_state(states, 1)
        assert np.array_equal(env.pack(nxt), orc.move(ost, 1)) and tc == [1.0] * 200
        prev = env.prev_state(nxt, 1)
        assert np.array_equal(env.pack(prev), orc.prev(orc.move(ost, 1), 1))
        if name in ("cube3", "lightsout7"):      # (n-puzzle: an illegal move is a no-op, so it has no inverse)
            assert np.array_equal(env.pack(prev), ost)
        # default template methods of the ABC agree with the fused override
        from deepcubea_b200.environments.environment_abstract import Environment
        exp2, _ = Environment.expand(env, states[:20])
        assert np.array_equal(np.stack([env.pack(r) for r in exp2]), och[:20])
        assert is_valid_soln(env.generate_goal_states(1)[0], [], env)


def _misplaced_fn(env):
    """Reference-style heuristic callable (numpy in / numpy out) -- exercises the adapter path of AStar."""
    goal_in = env.state_to_nnet_input(env.generate_goal_states(1))[0][0]

    def fn(states, is_nnet_format=False):
        x = states[0] if is_nnet_format else env.state_to_nnet_input(states)[0]
        return (x != goal_in[None]).sum(axis=1).astype(np.float64) / 8.0
    return fn


def test_astar_class_api_python_semantics():
    import random
    from deepcubea_b200.search_methods.astar import AStar, get_path
    from deepcubea_b200.utils.env_utils import get_environment
    from deepcubea_b200.utils.search_utils import is_valid_soln
    env = get_environment("cube3")
    np.random.seed(8); random.seed(8)
    states, _ = env.generate_states(3, (3, 6))
    h = _misplaced_fn(env)
    astar = AStar(states, env, h, [0.8] * 3, max_nodes=1 << 20)
    steps = 0
    while not min(astar.has_found_goal()):
        astar.step(h, 100)
        steps += 1
        assert steps < 200
    for i, s in enumerate(states):
        node = astar.get_goal_node_smallest_path_cost(i)
        path, soln, cost = get_path(node)
        assert node.is_solved and cost == len(soln) and path[0] == s
        assert is_valid_soln(s, soln, env)
        assert astar.get_num_nodes_generated(i) > 0
    assert len(astar.get_popped_nodes()) == 3


@pytest.mark.skipif(not os.path.exists(os.path.join(WEIGHTS_DIR, "model_state_dict.pt")), reason="trained weights not present (assets/)")
@pytest.mark.parametrize("language,precision", [("cuda", None), ("python", "fp32")])
def test_cli_solves_reference_test_states(tmp_path, golden_dir, language, precision):
    """`python search_methods/astar.py ...` on the first states of data/cube3/test: valid solutions, lengths within the
    reference's own optimality gap (<= optimal + 4, results/cube3), reference results.pkl schema."""
    sys.path.insert(0, ROOT)
    from environments.cube3 import Cube3State
    g = np.load(golden_dir + "/optimal_cube3.npz")
    n = 4
    states = [Cube3State(g["states"][i].astype(np.int64)) for i in range(n)]       # shipped pickles hold int64 payloads
    opt = np.diff(g["offsets"])[:n]
    pickle.dump({"states": states}, open(tmp_path / "in.pkl", "wb"))
    out = tmp_path / "res"
    cmd = [sys.executable, os.path.join(ROOT, "search_methods", "astar.py"), "--states", str(tmp_path / "in.pkl"), "--model", WEIGHTS_DIR,
           "--env", "cube3", "--weight", "0.6", "--batch_size", "2000", "--results_dir", str(out), "--language", language,
           "--nnet_batch_size", "10000", "--max_nodes", str(1 << 24)] + (["--nnet_precision", precision] if precision else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    res = pickle.load(open(out / "results.pkl", "rb"))
    assert sorted(res.keys()) == ["num_nodes_generated", "paths", "solutions", "states", "times"]
    from deepcubea_b200.utils.env_utils import get_environment
    from deepcubea_b200.utils.search_utils import is_valid_soln
    env = get_environment("cube3")
    for i in range(n):
        assert is_valid_soln(states[i], res["solutions"][i], env)
        assert len(res["paths"][i]) == len(res["solutions"][i]) + 1
        assert opt[i] <= len(res["solutions"][i]) <= opt[i] + 4
        assert res["num_nodes_generated"][i] > 0 and res["times"][i] > 0
    text = open(out / "output.txt").read()
    assert text.count("State: ") == n and "# Nodes Gen:" in text
    if precision is None:           # the reference's command line, no extra flag: the hand-written tcgen05 layers are the default
        import re
        m = re.search(r"nnet: precision=fp16x3 kernel=dcb_resnet_gemm \(tcgen05\) launches=(\d+)", text)
        assert m and int(m.group(1)) > 0, text[-500:]
    cmp = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "compare_solutions.py"), "--soln1", str(out / "results.pkl"),
                          "--soln2", str(out / "results.pkl")], capture_output=True, text=True, cwd=ROOT)
    assert cmp.returncode == 0 and "100.00% soln2 equal to soln1" in cmp.stdout


def test_bellman_backup_matches_oracle():
    """search_utils.bellman (utils/search_utils.py:16-32): expand + heuristic + min backup on the GPU env."""
    import random
    from deepcubea_b200.utils.env_utils import get_environment
    from deepcubea_b200.utils.search_utils import bellman
    from oracle import oracle_env as O
    from oracle.oracle_bwas import misplaced_heuristic
    for name in ("cube3", "puzzle24"):
        env, orc = get_environment(name), O.get_oracle_env(name)
        np.random.seed(12); random.seed(12)
        ost, _ = orc.generate_states(300, (0, 5))
        states = env.unpack(ost)
        h_np = misplaced_heuristic(orc)
        backup, per_state, exp = bellman(states, lambda sts: h_np(env.pack(sts)).astype(np.float64), env)
        och, _ = orc.expand(ost)
        ref = (1.0 + h_np(och.reshape(-1, orc.state_dim)).astype(np.float64)).reshape(300, orc.num_moves)
        assert np.allclose(np.stack(per_state), ref)
        assert np.allclose(backup, ref.min(axis=1) * ~orc.is_solved(ost))
        assert len(exp) == 300 and len(exp[0]) == orc.num_moves
        # the tensor-native form: same numbers, children / flags bit-exact, nothing leaves HBM until asked
        from deepcubea_b200.utils.search_utils import bellman_packed
        goal_in = torch.from_numpy(orc.nnet_input(orc.goal[None])[0]).cuda()
        b2, per2, ch2, sv2 = bellman_packed(ost, lambda x: (x != goal_in[None]).sum(dim=1).float() / 8.0, env)
        assert np.array_equal(ch2.cpu().numpy(), och)
        assert np.array_equal(sv2.cpu().numpy().astype(bool).reshape(-1), orc.is_solved(och.reshape(-1, orc.state_dim)))
        assert np.allclose(per2.cpu().numpy(), ref) and np.allclose(b2.cpu().numpy(), backup)


@pytest.mark.skipif(not os.path.exists(os.path.join(WEIGHTS_DIR, "model_state_dict.pt")), reason="trained weights not present (assets/)")
def test_cli_two_ranks_shard_the_states(tmp_path, golden_dir):
    """The CLI under torchrun, one process per GPU: start states sharded over the ranks, every rank's network and engine on ITS
    device, rank 0 alone owns output.txt / results.pkl.  Needs two GPUs (run with `gpurun --gpus 2`)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    from environments.cube3 import Cube3State
    g = np.load(golden_dir + "/optimal_cube3.npz")
    n = 6
    states = [Cube3State(g["states"][i].astype(np.int64)) for i in range(n)]
    opt = np.diff(g["offsets"])[:n]
    pickle.dump({"states": states}, open(tmp_path / "in.pkl", "wb"))
    out = tmp_path / "res"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29531",
           os.path.join(ROOT, "search_methods", "astar.py"), "--states", str(tmp_path / "in.pkl"), "--model", WEIGHTS_DIR, "--env", "cube3",
           "--weight", "0.6", "--batch_size", "2000", "--results_dir", str(out), "--max_nodes", str(1 << 26)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    res = pickle.load(open(out / "results.pkl", "rb"))
    from deepcubea_b200.utils.env_utils import get_environment
    from deepcubea_b200.utils.search_utils import is_valid_soln
    env = get_environment("cube3")
    assert len(res["solutions"]) == n
    for i in range(n):
        assert is_valid_soln(states[i], res["solutions"][i], env)
        assert opt[i] <= len(res["solutions"][i]) <= opt[i] + 4
    text = open(out / "output.txt").read()
    assert text.count("State: ") == n
    assert os.path.exists(out / "output.rank1.txt")
