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


def bellman_packed(states, heuristic_device_fn: Callable, env: Environment):
    """The same backup without leaving HBM and without a Python object per child -- the tensor-native form for the consumers that
    call `bellman` on millions of states (the reference's AVI target generation, updaters/updater.py, and gbfs.py).

    states: packed uint8 [n, S] (numpy or CUDA tensor).  heuristic_device_fn: nnet-input u8 [m, S] on the device -> f32 [m] on the
    device (`heuristic_fn.device_fn` of nnet_utils.load_heuristic_fn, e.g. the tcgen05 network).
    Returns CUDA tensors (backup f32 [n], per-action values f32 [n, A], children u8 [n, A, S], solved children u8 [n, A]).
    One dcb_expand launch produces children + their solved flags; transition costs are 1 (cube3.py:160, n_puzzle.py:171)."""
    import torch

    from .. import ops
    from .._lib import ENV_IDS
    from ..search_methods.astar import _env_name
    eid = ENV_IDS[_env_name(env)]
    st = states if isinstance(states, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(states, dtype=np.uint8))
    st = st.cuda().contiguous()
    n = st.shape[0]
    children, child_solved, _ = ops.expand(eid, st, want_hash=False)
    a, s = children.shape[1], children.shape[2]
    flat = children.reshape(n * a, s)
    h = heuristic_device_fn(ops.nnet_input(eid, flat)).float().clamp_(min=0.0)      # clip_zero (nnet_utils.py:193-194)
    per_action = 1.0 + h.reshape(n, a)
    unsolved = ops.is_solved(eid, st) == 0
    backup = per_action.min(dim=1).values * unsolved
    return backup, per_action, children, child_solved
