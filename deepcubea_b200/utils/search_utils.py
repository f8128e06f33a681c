"""Solution replay and one-step Bellman backup -- API of the reference's utils/search_utils.py:7-32."""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np

from ..environments.environment_abstract import Environment, State


def is_valid_soln(state: State, soln: Sequence[int], env: Environment) -> bool:
    """True iff applying `soln` move by move to `state` ends in the goal (the assert behind every reported solve,
    search_methods/astar.py:443, 556)."""
    frontier: List[State] = [state]
    for action in soln:
        frontier, _ = env.next_state(frontier, int(action))
    return bool(env.is_solved(frontier)[0])


def bellman(states: List[State], heuristic_fn: Callable, env: Environment) -> Tuple[np.ndarray, List[np.ndarray], List[List[State]]]:
    """backup(s) = 0 if s is solved else min_a [ cost(s,a) + h(child(s,a)) ].

    Returns (backup per state, the bracketed term per state and action, the children)."""
    children, step_costs = env.expand(states)
    widths = [len(row) for row in children]
    child_values = np.asarray(heuristic_fn([c for row in children for c in row]), dtype=np.float64)
    lookahead = np.concatenate(step_costs, axis=0) + child_values
    bounds = np.cumsum(widths)[:-1]
    per_state = np.split(lookahead, bounds)
    unsolved = np.logical_not(env.is_solved(states))
    backup = np.fromiter((row.min() for row in per_state), dtype=np.float64, count=len(per_state)) * unsolved
    return backup, per_state, children
