"""GPU parity: the HBM-resident BWAS engine vs the sequential oracle, trace-exact with an exactly
representable heuristic (same popped nodes, same kept nodes, same node ids, same solution)."""
import random

import numpy as np
import pytest

from oracle import oracle_env as O
from oracle.oracle_bwas import bwas, misplaced_heuristic

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _torch_misplaced(env):
    goal_in = torch.from_numpy(env.nnet_input(env.goal[None])[0]).cuda()

    def h(x):   # x: nnet-input u8 [m, S] on device
        return (x != goal_in[None]).sum(dim=1).to(torch.float32) / 8.0
    return h


@pytest.mark.parametrize("name,back,batch", [("cube3", (4, 9), 7), ("cube3", (6, 11), 100), ("cube3", (5, 8), 1),
                                              ("puzzle15", (10, 30), 50), ("puzzle24", (10, 30), 33),
                                              ("puzzle35", (10, 30), 64), ("puzzle48", (10, 40), 20), ("lightsout7", (3, 6), 25),
                                              ("cube4", (3, 6), 20)])
def test_engine_matches_oracle_trace(name, back, batch):
    from deepcubea_b200.search.bwas_gpu import BWASGpu
    env = O.get_oracle_env(name)
    np.random.seed(11); random.seed(11)
    states, _ = env.generate_states(6, back)
    eng = BWASGpu(name, _torch_misplaced(env), 0.8, batch, max_nodes=1 << 20)
    for s in states:
        ref = bwas(env, s, misplaced_heuristic(env), 0.8, batch, batch_dedup="min", keep_trace=True, max_iters=400)
        got = eng.solve(s, keep_trace=True, max_iters=400)
        assert got.iterations == ref["iterations"]
        for it, (a, b) in enumerate(zip(got.trace, ref["trace"])):
            assert a["popped"] == b["popped"], "iteration %d popped differ" % it
            assert a["kept"] == sorted(b["kept"]), "iteration %d kept differ" % it
        assert got.nodes_generated == ref["nodes_generated"]
        assert got.moves == ref["moves"]
        if got.moves is not None:
            cur = s[None]
            for mv in got.moves:
                cur = env.move(cur, mv)
            assert env.is_solved(cur)[0]                      # search_utils.is_valid_soln


def test_solved_start_and_empty_solution():
    from deepcubea_b200.search.bwas_gpu import BWASGpu
    env = O.get_oracle_env("cube3")
    eng = BWASGpu("cube3", _torch_misplaced(env), 0.8, 100, max_nodes=1 << 18)
    got = eng.solve(env.goal)
    ref = bwas(env, env.goal, misplaced_heuristic(env), 0.8, 100)
    assert got.moves == [] == ref["moves"]
    assert got.nodes_generated == ref["nodes_generated"]


@pytest.mark.parametrize("name,back,batch,weight", [("cube3", (4, 8), 10, 1.0), ("cube3", (5, 9), 100, 0.5), ("cube3", (3, 6), 1, 1.0),
                                                     ("puzzle15", (10, 24), 7, 0.5), ("puzzle48", (10, 30), 33, 1.0), ("lightsout7", (3, 5), 10, 0.5),
                                                     ("cube4", (3, 5), 10, 1.0)])
def test_engine_python_semantics_matches_oracle_trace(name, back, batch, weight):
    """semantics="python" (the `AStar` class path, astar.py:232-340) trace-exact against oracle.bwas_python (itself pinned to
    the reference's Python AStar)."""
    from deepcubea_b200.search.bwas_gpu import BWASGpu
    from oracle.oracle_bwas import bwas_python
    env = O.get_oracle_env(name)
    np.random.seed(17); random.seed(17)
    states, _ = env.generate_states(5, back)
    eng = BWASGpu(name, _torch_misplaced(env), weight, batch, max_nodes=1 << 20, semantics="python")
    for s in states:
        ref = bwas_python(env, s, misplaced_heuristic(env), weight, batch, batch_dedup="min", keep_trace=True, cost_dtype=np.float32)
        eng.reset(s)
        trace = []
        while not eng.goal_ids and eng.iterations < 3000:
            trace.append(eng.step(keep_trace=True))
        assert eng.iterations == ref["steps"]
        for it, (a, b) in enumerate(zip(trace, ref["trace"])):
            assert a["popped"] == b["popped"], "step %d popped differ" % it
            assert a["kept"] == sorted(b["kept"]), "step %d kept differ" % it
        assert eng.nodes_generated == ref["nodes_generated"]
        assert eng.goal_id == ref["goal_id"]
        assert eng.path_to(eng.goal_id) == ref["moves"]
