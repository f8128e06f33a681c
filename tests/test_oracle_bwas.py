"""CPU tier: the oracle's BWAS restatement vs the REAL reference binary (oracle/_ref, compiled from the
reference's cpp/*.cpp) over the reference's socket protocol, with a tie-free heuristic so that the C++ heap's
unspecified tie order cannot matter: moves, nodes generated and iteration count must be identical."""
import random

import numpy as np
import pytest

from oracle import oracle_env as O
from oracle.oracle_bwas import bwas, misplaced_heuristic
from oracle.ref_runner import HeuristicServer, have_reference_binary, run_reference_bwas


def _tie_free(env):
    base = misplaced_heuristic(env)
    wj = (np.arange(env.state_dim) + 1).astype(np.int64)

    def h(states):
        pert = ((states.astype(np.int64) * wj[None]).sum(axis=1) % 1009).astype(np.float32) / np.float32(65536.0)
        return (base(states) + pert).astype(np.float32)
    return h


@pytest.mark.skipif(not have_reference_binary(), reason="oracle/_ref/parallel_weighted_astar not built (needs /root/reference)")
@pytest.mark.parametrize("name,back,batch,weight", [("cube3", (4, 8), 100, 0.8), ("cube3", (5, 9), 10, 0.6), ("cube3", (3, 6), 1, 1.0),
                                                     ("puzzle15", (10, 24), 20, 0.8), ("puzzle48", (8, 16), 50, 0.6), ("cube4", (3, 6), 20, 0.8)])
def test_oracle_bwas_equals_reference_binary(name, back, batch, weight):
    env = O.get_oracle_env(name)
    h = _tie_free(env)
    np.random.seed(21); random.seed(21)
    states, _ = env.generate_states(4, back)
    srv = HeuristicServer(env.state_dim, h)
    try:
        for s in states:
            ref = run_reference_bwas(name, s, weight, batch, srv, timeout=120)
            ours = bwas(env, s, h, weight, batch, batch_dedup="sequential")
            assert ours["moves"] == ref["moves"]
            assert ours["nodes_generated"] == ref["nodes_generated"]
            assert ours["iterations"] == ref["iterations"]
            cur = s[None]
            for mv in ref["moves"]:
                cur = env.move(cur, mv)
            assert env.is_solved(cur)[0]
    finally:
        srv.close()


@pytest.mark.parametrize("name", ["cube3", "puzzle15"])
def test_min_and_sequential_dedup_agree_on_validity(name):
    """The GPU's in-batch duplicate rule ("min") vs the reference's child-order rule: both return valid
    solutions; lengths agree on these cases (the rule only removes dominated duplicates)."""
    env = O.get_oracle_env(name)
    h = misplaced_heuristic(env)
    np.random.seed(5); random.seed(5)
    states, _ = env.generate_states(5, (4, 9) if name == "cube3" else (10, 30))
    for s in states:
        a = bwas(env, s, h, 0.8, 50, batch_dedup="min")
        b = bwas(env, s, h, 0.8, 50, batch_dedup="sequential")
        for r in (a, b):
            cur = s[None]
            for mv in r["moves"]:
                cur = env.move(cur, mv)
            assert env.is_solved(cur)[0]
        assert len(a["moves"]) == len(b["moves"])
        assert a["nodes_generated"] <= b["nodes_generated"]


def test_oracle_python_semantics_equals_reference_astar(golden_dir):
    """oracle.bwas_python vs traces of the reference's own Python `AStar` (tests/golden/make_golden_astar.py): same moves,
    nodes generated, step count, pops per step, OPEN / CLOSED sizes on 60 cube3 + puzzle15 cases; weights 0.8 / 0.6 pin the float64
    cost arithmetic of astar.py:196."""
    import json
    from oracle.oracle_bwas import bwas_python
    cases = json.load(open(golden_dir + "/astar_python_traces.json"))
    assert len(cases) == 60
    for c in cases:
        env = O.get_oracle_env(c["env"])
        r = bwas_python(env, np.array(c["state"], np.uint8), misplaced_heuristic(env), c["weight"], c["batch"], batch_dedup="sequential")
        assert r["moves"] == c["moves"], c
        assert r["nodes_generated"] == c["nodes_generated"] and r["steps"] == c["steps"]
        assert r["popped_per_step"] == c["popped_per_step"]
        assert r["open_size"] == c["open_size"] and r["closed_size"] == c["closed_size"]
        if c["weight"] in (1.0, 0.5):
            r32 = bwas_python(env, np.array(c["state"], np.uint8), misplaced_heuristic(env), c["weight"], c["batch"], cost_dtype=np.float32)
            assert r32["moves"] == c["moves"] and r32["nodes_generated"] == c["nodes_generated"]     # w*g exact: fp32 == fp64


def _inconsistent(env, amp):
    base = misplaced_heuristic(env)
    wj = (np.arange(env.state_dim) * 2654435761 % 1000003 + 1).astype(np.int64)

    def h(states):
        r = ((states.astype(np.int64) * wj[None]).sum(axis=1) * 2654435761 % 1000003).astype(np.float32) / np.float32(1000003.0)
        return (base(states) * np.float32(1.5) + r * np.float32(amp)).astype(np.float32)
    return h


@pytest.mark.parametrize("name,back,batch,weight,amp", [("cube3", (6, 11), 200, 0.6, 4.0), ("puzzle15", (15, 40), 50, 0.8, 4.0)])
def test_stored_node_rewrite_changes_nothing(name, back, batch, weight, amp):
    """parallel_weighted_astar.cpp:255-257 rewrites depth / parent of the node object stored in CLOSED when its state is re-reached
    more cheaply.  The GPU engine leaves the older node alone.  Under an inconsistent heuristic that triggers
    hundreds of rewrites the search result is the same either way (tools/exp_stored_node_rewrite.py: 60/60 over 8,514 rewrites)."""
    env = O.get_oracle_env(name)
    h = _inconsistent(env, amp)
    np.random.seed(3); random.seed(3)
    states, _ = env.generate_states(6, back)
    links = 0
    for s in states:
        a = bwas(env, s, h, weight, batch, mutate_stored=True, max_iters=150)
        b = bwas(env, s, h, weight, batch, mutate_stored=False, max_iters=150)
        links += a["links"]
        assert a["moves"] == b["moves"] and a["nodes_generated"] == b["nodes_generated"] and a["iterations"] == b["iterations"]
    assert links > 20, "the heuristic no longer provokes rewrites: the test proves nothing"
