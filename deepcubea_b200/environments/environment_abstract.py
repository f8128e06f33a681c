"""Plugin API kept verbatim from the reference (environments/environment_abstract.py:8-163): `State`
(hashable / comparable) and `Environment` with its seven abstract methods plus the `generate_states` and
`expand` template methods.  Concrete environments in this package run every batched operation on the GPU
through the C ABI (deepcubea_b200.ops); there is no CPU implementation behind them.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from random import randrange
from typing import List, Tuple

import numpy as np
import torch.nn as nn


class State(ABC):
    @abstractmethod
    def __hash__(self):
        pass

    @abstractmethod
    def __eq__(self, other):
        pass


class Environment(ABC):
    def __init__(self):
        self.dtype = float
        self.fixed_actions: bool = True

    @abstractmethod
    def next_state(self, states: List[State], action: int) -> Tuple[List[State], List[float]]:
        """Next states and transition costs for one action applied to every state."""

    @abstractmethod
    def prev_state(self, states: List[State], action: int) -> List[State]:
        """States from which `action` leads to the given states."""

    @abstractmethod
    def generate_goal_states(self, num_states: int) -> List[State]:
        pass

    @abstractmethod
    def is_solved(self, states: List[State]) -> np.ndarray:
        """Boolean array, element i tells whether states[i] is the goal."""

    @abstractmethod
    def state_to_nnet_input(self, states: List[State]) -> List[np.ndarray]:
        pass

    @abstractmethod
    def get_num_moves(self) -> int:
        pass

    @abstractmethod
    def get_nnet_model(self) -> nn.Module:
        pass

    def generate_states(self, num_states: int, backwards_range: Tuple[int, int]) -> Tuple[List[State], List[int]]:
        """Scramble goal states with random reverse moves; each state gets a depth drawn uniformly from
        [backwards_range[0], backwards_range[1]].  Draws from numpy's and `random`'s global generators in the order the
        reference does (environment_abstract.py:88-125), so a fixed seed yields the reference's states."""
        lo, hi = backwards_range
        if num_states <= 0 or lo < 0:
            raise AssertionError("num_states must be positive and the scramble range non-negative")
        assert self.fixed_actions, "Environments without fixed actions must implement their own method"
        n_actions = self.get_num_moves()
        population: List[State] = self.generate_goal_states(num_states)
        target_depth = np.random.choice(np.arange(lo, hi + 1).tolist(), num_states)
        applied = np.zeros(num_states)
        pending = applied < target_depth
        while pending.max():
            candidates = np.flatnonzero(pending)
            picked = np.random.choice(candidates, int(max(len(candidates) / n_actions, 1)))
            action = randrange(n_actions)
            for slot, moved in zip(picked, self.prev_state([population[i] for i in picked], action)):
                population[slot] = moved
            applied[picked] = applied[picked] + 1
            pending = applied < target_depth
        return population, target_depth.tolist()

    def expand(self, states: List[State]) -> Tuple[List[List[State]], List[np.ndarray]]:
        """Children of every state for every action (action-minor) and the matching transition costs
        (environment_abstract.py:127-163).  Concrete environments override this with one fused GPU launch."""
        assert self.fixed_actions, "Environments without fixed actions must implement their own method"
        per_action = [self.next_state(states, action) for action in range(self.get_num_moves())]
        children = [[nxt[i] for nxt, _ in per_action] for i in range(len(states))]
        costs = np.array([tc for _, tc in per_action], dtype=np.float64).T.reshape(len(states), len(per_action))
        return children, [costs[i] for i in range(len(states))]
