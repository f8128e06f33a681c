"""Solution validation and Bellman backup (utils/search_utils.py:7-32)."""
from typing import List, Tuple

import numpy as np

from ..environments.environment_abstract import Environment, State
from . import misc_utils


def is_valid_soln(state: State, soln: List[int], env: Environment) -> bool:
    """Replay the moves from `state` and check the result is the goal."""
    cur = state
    for move in soln:
        cur = env.next_state([cur], move)[0][0]
    return bool(env.is_solved([cur])[0])


def bellman(states: List, heuristic_fn, env: Environment) -> Tuple[np.ndarray, List[np.ndarray], List[List[State]]]:
    """One-step lookahead backup: min_a (tc + h(child)), zero for solved states."""
    states_exp, tc_l = env.expand(states)
    tc = np.concatenate(tc_l, axis=0)
    flat, split_idxs = misc_utils.flatten(states_exp)
    ctg_next_p_tc = tc + heuristic_fn(flat)
    per_state = np.split(ctg_next_p_tc, split_idxs)
    backup = np.array([np.min(x) for x in per_state]) * np.logical_not(env.is_solved(states))
    return backup, per_state, states_exp
