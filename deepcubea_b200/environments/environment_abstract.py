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
        """Scramble from the goal by random reverse moves (environment_abstract.py:88-125); the numpy / `random`
        call sequence is the reference's, so fixed seeds give the reference's states."""
        assert num_states > 0 and backwards_range[0] >= 0
        assert self.fixed_actions, "Environments without fixed actions must implement their own method"
        depths = list(range(backwards_range[0], backwards_range[1] + 1))
        n_moves = self.get_num_moves()
        states: List[State] = self.generate_goal_states(num_states)
        scramble_nums = np.random.choice(depths, num_states)
        done_moves = np.zeros(num_states)
        while np.max(done_moves < scramble_nums):
            idxs = np.where(done_moves < scramble_nums)[0]
            idxs = np.random.choice(idxs, int(max(len(idxs) / n_moves, 1)))
            move = randrange(n_moves)
            moved = self.prev_state([states[i] for i in idxs], move)
            for k, s in enumerate(moved):
                states[idxs[k]] = s
            done_moves[idxs] = done_moves[idxs] + 1
        return states, scramble_nums.tolist()

    def expand(self, states: List[State]) -> Tuple[List[List[State]], List[np.ndarray]]:
        """All children of every state, move-minor (environment_abstract.py:127-163)."""
        assert self.fixed_actions, "Environments without fixed actions must implement their own method"
        n, n_moves = len(states), self.get_num_moves()
        children: List[List[State]] = [[] for _ in range(n)]
        tc = np.empty([n, n_moves])
        for move in range(n_moves):
            nxt, tc_move = self.next_state(states, move)
            tc[:, move] = np.array(tc_move)
            for i in range(n):
                children[i].append(nxt[i])
        return children, [tc[i] for i in range(n)]
