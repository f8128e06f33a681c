"""GPU parity: the HBM-resident BWAS engine vs the sequential oracle, trace-exact with an exactly
representable heuristic (same popped nodes, same kept nodes, same node ids, same solution) -- the oracle runs the reference's
child-order CLOSED loop (`batch_dedup="sequential"`), the variant pinned to the reference binary and to the reference's Python
AStar traces (tests/test_oracle_bwas.py); and the engine head to head against the reference binary itself."""
import random

import numpy as np
import pytest

from oracle import oracle_env as O
from oracle.oracle_bwas import bwas, bwas_python, misplaced_heuristic

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _torch_misplaced(env):
    goal_in = torch.from_numpy(env.nnet_input(env.goal[None])[0]).cuda()

    def h(x):   # x: nnet-input u8 [m, S] on device
        return (x != goal_in[None]).sum(dim=1).to(torch.float32) / 8.0
    return h


def _noisy_np(env, amp=3.0):
    """A deliberately INCONSISTENT heuristic on the nnet input (exact float32 arithmetic on both sides): states get re-reached with
    smaller depth, several depths of one state meet inside one batch -- the cases where the CLOSED rules differ."""
    goal_in = env.nnet_input(env.goal[None])[0]
    wj = (np.arange(env.state_dim, dtype=np.int64) * 7919 % 1009 + 1)

    def h(states):
        x = env.nnet_input(states)
        base = (x != goal_in[None]).sum(axis=1).astype(np.float32)
        r = ((x.astype(np.int64) * wj[None]).sum(axis=1) * 31 % 1024).astype(np.float32)
        return (base * np.float32(0.1875) + r * np.float32(amp / 1024.0)).astype(np.float32)
    return h


def _noisy_torch(env, amp=3.0):
    goal_in = torch.from_numpy(env.nnet_input(env.goal[None])[0]).cuda()
    wj = torch.from_numpy(np.arange(env.state_dim, dtype=np.int64) * 7919 % 1009 + 1).cuda()

    def h(x):
        base = (x != goal_in[None]).sum(dim=1).to(torch.float32)
        r = ((x.to(torch.int64) * wj[None]).sum(dim=1) * 31 % 1024).to(torch.float32)
        return base * 0.1875 + r * (amp / 1024.0)
    return h


def _check_trace(got, ref):
    assert got.iterations == ref["iterations"]
    for it, (a, b) in enumerate(zip(got.trace, ref["trace"])):
        assert a["popped"] == b["popped"], "iteration %d popped differ" % it
        assert a["kept"] == sorted(b["kept"]), "iteration %d kept differ" % it
    assert got.nodes_generated == ref["nodes_generated"]
    assert got.moves == ref["moves"]


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
        ref = bwas(env, s, misplaced_heuristic(env), 0.8, batch, keep_trace=True, max_iters=400)
        got = eng.solve(s, keep_trace=True, max_iters=400)
        _check_trace(got, ref)
        if got.moves is not None:
            cur = s[None]
            for mv in got.moves:
                cur = env.move(cur, mv)
            assert env.is_solved(cur)[0]                      # search_utils.is_valid_soln


@pytest.mark.parametrize("name,back,batch,weight,amp", [("cube3", (5, 9), 50, 0.8, 3.0), ("cube3", (5, 9), 20, 0.2, 3.0),
                                                         ("puzzle15", (15, 40), 100, 0.3, 6.0), ("lightsout7", (3, 6), 30, 0.4, 3.0)])
def test_engine_matches_oracle_trace_inconsistent_heuristic(name, back, batch, weight, amp):
    """The reference's CLOSED loop is sequential in child order (parallel_weighted_astar.cpp:246-261): with an inconsistent
    heuristic one batch holds the same state at several depths, the deeper one first.  Trace-exact all the same."""
    from deepcubea_b200.search.bwas_gpu import BWASGpu
    env = O.get_oracle_env(name)
    np.random.seed(3); random.seed(3)
    states, _ = env.generate_states(6, back)
    eng = BWASGpu(name, _noisy_torch(env, amp), weight, batch, max_nodes=1 << 21)
    for s in states:
        ref = bwas(env, s, _noisy_np(env, amp), weight, batch, keep_trace=True, max_iters=300)
        got = eng.solve(s, keep_trace=True, max_iters=300)
        _check_trace(got, ref)


@pytest.mark.parametrize("name,back,batch", [("cube3", (5, 9), 64), ("puzzle15", (12, 30), 40), ("lightsout7", (3, 6), 25)])
def test_pipelined_sync_free_solve_equals_stepwise(name, back, batch):
    """solve() without host round trips inside the iteration (device-side counts, one iteration in flight ahead of the host)
    returns exactly what the step-by-step run returns."""
    from deepcubea_b200.search.bwas_gpu import BWASGpu
    env = O.get_oracle_env(name)
    np.random.seed(23); random.seed(23)
    states, _ = env.generate_states(5, back)
    eng = BWASGpu(name, _torch_misplaced(env), 0.8, batch, max_nodes=1 << 20, sync_free=True)
    assert eng.sync_free
    for s in states:
        ref = bwas(env, s, misplaced_heuristic(env), 0.8, batch)
        got = eng.solve(s)
        assert got.moves == ref["moves"] and got.nodes_generated == ref["nodes_generated"] and got.iterations == ref["iterations"]
        assert got.open_size == ref["open_size"] and got.closed_size == ref["closed_size"]


def test_solved_start_and_empty_solution():
    from deepcubea_b200.search.bwas_gpu import BWASGpu
    env = O.get_oracle_env("cube3")
    eng = BWASGpu("cube3", _torch_misplaced(env), 0.8, 100, max_nodes=1 << 18)
    got = eng.solve(env.goal)
    ref = bwas(env, env.goal, misplaced_heuristic(env), 0.8, 100)
    assert got.moves == [] == ref["moves"]
    assert got.nodes_generated == ref["nodes_generated"]


def test_closed_table_grows_with_the_search():
    """CLOSED starts small (L2-resident) and is rehashed into larger tables as the search fills it; results do not change."""
    from deepcubea_b200.search.bwas_gpu import BWASGpu
    env = O.get_oracle_env("cube3")
    np.random.seed(4); random.seed(4)
    states, _ = env.generate_states(3, (9, 11))
    eng = BWASGpu("cube3", _torch_misplaced(env), 0.8, 200, max_nodes=1 << 22)
    small = eng.closed_cap_min
    grew = 0
    for s in states:
        ref = bwas(env, s, misplaced_heuristic(env), 0.8, 200, max_iters=60)
        got = eng.solve(s, max_iters=60)
        assert got.nodes_generated == ref["nodes_generated"] and got.moves == ref["moves"] and got.closed_size == ref["closed_size"]
        grew += eng.closed_cap > small
    assert grew > 0, "the searches never outgrew the initial table: make them longer"
    eng.reset(states[0])
    assert eng.closed_cap == small           # a new search starts from the small table again


@pytest.mark.parametrize("name,back,batch,weight", [("cube3", (4, 8), 10, 1.0), ("cube3", (5, 9), 100, 0.5), ("cube3", (3, 6), 1, 1.0),
                                                     ("puzzle15", (10, 24), 7, 0.5), ("puzzle48", (10, 30), 33, 1.0), ("lightsout7", (3, 5), 10, 0.5),
                                                     ("cube4", (3, 5), 10, 1.0), ("cube3", (5, 9), 50, 0.8), ("cube3", (5, 9), 20, 0.6),
                                                     ("puzzle15", (10, 24), 30, 0.8), ("lightsout7", (3, 5), 10, 0.3)])
def test_engine_python_semantics_matches_oracle_trace(name, back, batch, weight):
    """semantics="python" (the `AStar` class path, astar.py:232-340) trace-exact against oracle.bwas_python (itself pinned to
    the reference's Python AStar), costs in float64 like astar.py:196 (weights 0.8 / 0.6 / 0.3: w*g is not exact in float32)."""
    from deepcubea_b200.search.bwas_gpu import BWASGpu
    env = O.get_oracle_env(name)
    np.random.seed(17); random.seed(17)
    states, _ = env.generate_states(5, back)
    eng = BWASGpu(name, _torch_misplaced(env), weight, batch, max_nodes=1 << 20, semantics="python")
    for s in states:
        ref = bwas_python(env, s, misplaced_heuristic(env), weight, batch, keep_trace=True)
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


@pytest.mark.parametrize("name,back,n_inst,batch,weight,sync_free", [("cube3", (3, 8), 64, 100, 0.5, False), ("cube3", (3, 8), 64, 100, 1.0, True),
                                                                      ("puzzle15", (8, 18), 40, 20, 0.8, True), ("lightsout7", (2, 5), 9, 16, 0.6, False),
                                                                      ("cube3", (1, 4), 1000, 4, 0.5, True)])
def test_multi_instance_engine_matches_oracle_per_instance(name, back, n_inst, batch, weight, sync_free):
    """Many instances in ONE engine (one arena, one CLOSED keyed per instance, segmented OPEN, one heuristic call per step):
    every instance pops / keeps exactly what the reference's Python AStar does for it alone (astar.py:256-317)."""
    from deepcubea_b200.search.engine import SearchEngine
    env = O.get_oracle_env(name)
    np.random.seed(29); random.seed(29)
    states, _ = env.generate_states(n_inst, back)
    states[n_inst // 2] = env.goal                       # one instance starts solved: its root is popped and is the goal (0 moves)
    calls = [0]
    h_t = _torch_misplaced(env)

    def counted(x):
        calls[0] += 1
        return h_t(x)
    eng = SearchEngine(name, counted, [weight] * n_inst, batch, n_inst=n_inst, max_nodes=1 << 26, semantics="python", sync_free=sync_free)
    eng.reset(states)
    refs = [bwas_python(env, s, misplaced_heuristic(env), weight, batch, keep_trace=True) for s in states]
    A, npi = eng.A, eng.nodes_per_inst
    local = lambda ids, i: [x - i * npi for x in ids]                 # global node id -> the instance's own numbering
    steps = 0
    calls[0] = 0
    while eng.running() and steps < 3000:
        eng.step_all()
        kept = np.array(eng.kept_list(), dtype=np.int64)
        for i, rec in enumerate(eng.inst):
            if rec.resting:
                assert steps >= refs[i]["steps"]
                continue
            tr = refs[i]["trace"][steps]
            assert local(eng.popped_of(i), i) == tr["popped"], "instance %d step %d popped differ" % (i, steps)
            mine = kept[(kept >= i * npi) & (kept < (i + 1) * npi)] - i * npi
            assert mine.tolist() == sorted(tr["kept"]), "instance %d step %d kept differ" % (i, steps)
        steps += 1
    assert calls[0] == steps                                            # ONE heuristic evaluation per step for all instances
    assert steps == max(r["steps"] for r in refs)
    for i, (rec, r) in enumerate(zip(eng.inst, refs)):
        assert rec.iterations == r["steps"] and rec.nodes_generated == r["nodes_generated"]
        assert rec.goal_id - i * npi == r["goal_id"]
        assert eng.path_to(rec.goal_id) == r["moves"]
        assert rec.open_size == r["open_size"]


def _tie_free_pair(env):
    """The tie-free heuristic of tests/test_oracle_bwas.py on the nnet input, as a numpy function (for the reference binary's
    socket) and a torch function (for the engine), bit-identical in float32."""
    goal_in = env.nnet_input(env.goal[None])[0]
    wj_np = (np.arange(env.state_dim) + 1).astype(np.int64)
    goal_t = torch.from_numpy(goal_in).cuda()
    wj_t = torch.from_numpy(wj_np).cuda()

    def h_np(states):
        x = env.nnet_input(states)
        base = (x != goal_in[None]).sum(axis=1).astype(np.float32) / np.float32(8.0)
        pert = ((x.astype(np.int64) * wj_np[None]).sum(axis=1) * 2654435761 % (1 << 20)).astype(np.float32) / np.float32(1 << 24)
        return (base + pert).astype(np.float32)

    def h_t(x):
        base = (x != goal_t[None]).sum(dim=1).to(torch.float32) / 8.0
        pert = ((x.to(torch.int64) * wj_t[None]).sum(dim=1) * 2654435761 % (1 << 20)).to(torch.float32) / float(1 << 24)
        return base + pert
    return h_np, h_t


@pytest.mark.parametrize("name,back,batch,weight", [("cube3", (5, 9), 100, 0.8), ("cube3", (5, 9), 10, 0.6), ("puzzle15", (10, 24), 20, 0.8),
                                                     ("puzzle48", (8, 16), 50, 0.6), ("cube4", (3, 6), 20, 0.8)])
def test_engine_equals_reference_binary(name, back, batch, weight):
    """Head to head with the UNMODIFIED reference program (oracle/_ref/parallel_weighted_astar, compiled from the reference's
    cpp/*.cpp): same moves, same nodes generated, same iteration count -- no oracle in between."""
    from deepcubea_b200.search.bwas_gpu import BWASGpu
    from oracle.ref_runner import HeuristicServer, have_reference_binary, run_reference_bwas
    if not have_reference_binary():
        pytest.skip("oracle/_ref/parallel_weighted_astar not built")
    env = O.get_oracle_env(name)
    h_np, h_t = _tie_free_pair(env)
    np.random.seed(21); random.seed(21)
    states, _ = env.generate_states(4, back)
    eng = BWASGpu(name, h_t, weight, batch, max_nodes=1 << 21)
    srv = HeuristicServer(env.state_dim, h_np)
    try:
        for s in states:
            ref = run_reference_bwas(name, s, weight, batch, srv, timeout=300)
            got = eng.solve(s)
            assert got.moves == ref["moves"]
            assert got.nodes_generated == ref["nodes_generated"]
            assert got.iterations == ref["iterations"]
    finally:
        srv.close()


def test_include_solved_keeps_stepping_instances_with_a_goal():
    """AStar.step(include_solved=True) (astar.py:263-265): instances that already popped a goal node keep searching; the reported
    goal stays the one of smallest path cost.  Without the flag they rest (their node count stops growing)."""
    from deepcubea_b200.search.engine import SearchEngine
    env = O.get_oracle_env("cube3")
    np.random.seed(31); random.seed(31)
    states, _ = env.generate_states(6, (2, 5))
    engines = []
    for flag in (False, True):
        eng = SearchEngine("cube3", _torch_misplaced(env), [1.0] * 6, 8, n_inst=6, max_nodes=1 << 22, semantics="python")
        eng.reset(states)
        for _ in range(40):
            eng.step_all(include_solved=flag)
        engines.append(eng)
    rest, cont = engines
    for i in range(6):
        a, b = rest.inst[i], cont.inst[i]
        assert a.n_goals >= 1 and b.n_goals >= a.n_goals
        assert b.nodes_generated > a.nodes_generated                     # kept expanding after its goal
        assert b.goal_key <= a.goal_key                                   # never a worse answer
        assert len(cont.path_to(b.goal_id)) == b.goal_key                 # unit move costs: path cost == number of moves
        cur = states[i][None]
        for mv in cont.path_to(b.goal_id):
            cur = env.move(cur, mv)
        assert env.is_solved(cur)[0]


def test_astar_class_replays_the_references_python_traces(golden_dir):
    """The `AStar` class of this repository against traces of the REFERENCE's own Python AStar (tests/golden/astar_python_traces.json,
    written by make_golden_astar.py running the unmodified reference): 60 cases, instances of one (env, weight, batch) group advanced
    TOGETHER in one engine -- moves, nodes generated, steps, pops per step, OPEN size equal per instance, CLOSED size equal in total.
    Weights 0.8 / 0.6 exercise the float64 cost keys."""
    import json
    from collections import defaultdict
    from deepcubea_b200.search_methods.astar import AStar, get_path
    from deepcubea_b200.utils.env_utils import get_environment
    cases = json.load(open(golden_dir + "/astar_python_traces.json"))
    groups = defaultdict(list)
    for c in cases:
        groups[(c["env"], c["weight"], c["batch"])].append(c)
    assert len(cases) == 60 and len(groups) == 12
    for (env_name, weight, batch), grp in groups.items():
        env = get_environment(env_name)
        goal_in = env.state_to_nnet_input(env.generate_goal_states(1))[0][0]

        def fn(states, is_nnet_format=False, env=env, goal_in=goal_in):
            x = states[0] if is_nnet_format else env.state_to_nnet_input(states)[0]
            return (x != goal_in[None]).sum(axis=1).astype(np.float64) / 8.0
        states = env.unpack(np.array([c["state"] for c in grp], dtype=np.uint8))
        astar = AStar(states, env, fn, [weight] * len(grp), max_nodes=1 << 23)
        popped_per_step = [[] for _ in grp]
        steps = 0
        while not min(astar.has_found_goal()):
            before = [len(p) for p in astar.popped_ids]
            astar.step(fn, batch)
            steps += 1
            for i, rec in enumerate(astar.engine.inst):
                if not rec.resting:
                    popped_per_step[i].append(len(astar.popped_ids[i]) - before[i])
            assert steps < 5000
        closed_total = 0
        for i, c in enumerate(grp):
            rec = astar.engine.inst[i]
            _, soln, cost = get_path(astar.get_goal_node_smallest_path_cost(i))
            assert [int(m) for m in soln] == c["moves"] and cost == c["path_cost"], (env_name, weight, batch, i)
            assert rec.iterations == c["steps"] and astar.get_num_nodes_generated(i) == c["nodes_generated"]
            assert popped_per_step[i] == c["popped_per_step"]
            assert rec.open_size == c["open_size"]
            closed_total += c["closed_size"]
        assert int(astar.engine.plan.closed_entries) == closed_total
