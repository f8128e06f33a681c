#!/usr/bin/env python
"""Golden traces of the reference's PYTHON batch weighted A* (search_methods/astar.py:232-340, the `AStar` class) with an
exactly representable heuristic, produced by running the UNMODIFIED reference here (build container only):

    python tests/golden/make_golden_astar.py      -> tests/golden/astar_python_traces.json

Heuristic = (#positions of the nnet input that differ from the goal's) / 8 -- the same function as
oracle.oracle_bwas.misplaced_heuristic; weights 1.0 / 0.5 keep w*g exact in float32 and float64 alike, 0.8 / 0.6 do not: those
cases pin the float64 cost arithmetic of astar.py:196.
Shims (no source edits): np.float / np.int for numpy 2; State.__hash__ via .tobytes() (numpy 2 has no .tostring()).
"""
import json
import os
import random
import sys

import numpy as np

np.float = float
np.int = int
REF = os.environ.get("DCB_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
OUT = os.path.dirname(os.path.abspath(__file__))

from environments.cube3 import Cube3, Cube3State  # noqa: E402
from environments.n_puzzle import NPuzzle, NPuzzleState  # noqa: E402
from search_methods.astar import AStar, get_path  # noqa: E402

Cube3State.__hash__ = lambda self: hash(np.asarray(self.colors).tobytes())
NPuzzleState.__hash__ = lambda self: hash(np.asarray(self.tiles).tobytes())


def main():
    cases = []
    for env_name, env, back in (("cube3", Cube3(), (3, 7)), ("puzzle15", NPuzzle(4), (8, 20))):
        goal_in = env.state_to_nnet_input(env.generate_goal_states(1))[0][0]

        def heuristic_fn(states, is_nnet_format=False, env=env, goal_in=goal_in):
            x = env.state_to_nnet_input(states)[0]
            return (x != goal_in[None]).sum(axis=1).astype(np.float64) / 8.0

        np.random.seed(31); random.seed(31)
        states, _ = env.generate_states(5, back)
        attr = "colors" if env_name == "cube3" else "tiles"
        for weight, batch in ((1.0, 1), (1.0, 10), (0.5, 100), (0.5, 7), (0.8, 10), (0.6, 100)):
            for s in states:
                astar = AStar([s], env, heuristic_fn, [weight])
                steps = 0
                popped_per_step = []
                while not min(astar.has_found_goal()):
                    before = len(astar.instances[0].popped_nodes)
                    astar.step(heuristic_fn, batch)
                    popped_per_step.append(len(astar.instances[0].popped_nodes) - before)
                    steps += 1
                    assert steps < 5000
                goal = astar.get_goal_node_smallest_path_cost(0)
                _, soln, cost = get_path(goal)
                cases.append({"env": env_name, "state": [int(v) for v in getattr(s, attr)], "weight": weight, "batch": batch,
                              "moves": [int(m) for m in soln], "path_cost": float(cost), "steps": steps,
                              "nodes_generated": int(astar.get_num_nodes_generated(0)), "popped_per_step": popped_per_step,
                              "closed_size": len(astar.instances[0].closed_dict), "open_size": len(astar.instances[0].open_set)})
    json.dump(cases, open(os.path.join(OUT, "astar_python_traces.json"), "w"))
    print(len(cases), "cases;", sum(c["nodes_generated"] for c in cases), "nodes")


if __name__ == "__main__":
    main()
