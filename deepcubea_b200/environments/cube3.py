"""3x3x3 Rubik's cube behind the reference's Environment API (environments/cube3.py:10-171), GPU-backed.

State = 54 sticker ids (uint8), goal = arange(54); 12 moves U-1 U1 D-1 D1 L-1 L1 R-1 R1 B-1 B1 F-1 F1.
All batched work (next_state / expand / is_solved / state_to_nnet_input) runs in hand-written CUDA through
the C ABI; the move permutations are built geometrically in cube3_geometry.py and compiled into the kernels.
`Cube3State` keeps the reference's module path and slots so the shipped pickles load and our results.pkl
unpickles inside the reference tree.
"""
from __future__ import annotations

from random import randrange
from typing import List, Tuple, Union

import numpy as np
from torch import nn

from .. import ops
from .._lib import ENV_IDS
from ..utils.pytorch_models import ResnetModel
from . import cube3_geometry
from ._packed import PackedEnvMixin
from .environment_abstract import Environment, State


class Cube3State(State):
    __slots__ = ["colors", "hash"]

    def __init__(self, colors: np.ndarray):
        self.colors: np.ndarray = colors
        self.hash = None

    def __hash__(self):
        # shipped pickles may leave the `hash` slot unset; numpy 2 has no .tostring() (cube3.py:17-21)
        h = getattr(self, "hash", None)
        if h is None:
            h = hash(np.asarray(self.colors).tobytes())
            self.hash = h
        return h

    def __eq__(self, other):
        return np.array_equal(self.colors, other.colors)


Cube3State.__module__ = "environments.cube3"     # pickle-compatible with the reference tree


class Cube3(PackedEnvMixin, Environment):
    moves: List[str] = cube3_geometry.MOVES
    moves_rev: List[str] = cube3_geometry.MOVES_REV
    env_id = ENV_IDS["cube3"]
    state_dim = 54
    _state_cls = Cube3State
    _attr = "colors"

    def __init__(self):
        super().__init__()
        self.dtype = np.uint8
        self.cube_len = 3
        self.goal_colors: np.ndarray = np.arange(0, 54, 1, dtype=self.dtype)
        self._rev_action = cube3_geometry.inverse_actions()

    def next_state(self, states: List[Cube3State], action: int) -> Tuple[List[Cube3State], List[float]]:
        nxt, tcs = self._next_state_np(self.pack(states), action)
        return self.unpack(nxt), tcs

    def prev_state(self, states: List[Cube3State], action: int) -> List[Cube3State]:
        return self.next_state(states, self._rev_action[action])[0]

    def generate_goal_states(self, num_states: int, np_format: bool = False) -> Union[List[Cube3State], np.ndarray]:
        if np_format:
            return np.repeat(self.goal_colors[None, :].copy(), num_states, axis=0)
        return [Cube3State(self.goal_colors.copy()) for _ in range(num_states)]

    def is_solved(self, states: List[Cube3State]) -> np.ndarray:
        return self._is_solved_np(self.pack(states))

    def state_to_nnet_input(self, states: List[Cube3State]) -> List[np.ndarray]:
        x = ops.nnet_input(self.env_id, self.to_device(self.pack(states))).cpu().numpy()
        return [x.astype(self.dtype, copy=False)]

    def get_num_moves(self) -> int:
        return len(self.moves)

    def get_nnet_model(self) -> nn.Module:
        return ResnetModel(54, 6, 5000, 1000, 4, 1, True)

    def generate_states(self, num_states: int, backwards_range: Tuple[int, int]) -> Tuple[List[Cube3State], List[int]]:
        """cube3.py:96-127 with the state array resident on the GPU (same RNG call sequence)."""
        assert num_states > 0 and backwards_range[0] >= 0
        import torch
        depths = list(range(backwards_range[0], backwards_range[1] + 1))
        st = self.to_device(self.generate_goal_states(num_states, np_format=True))
        scramble_nums = np.random.choice(depths, num_states)
        done_moves = np.zeros(num_states)
        lt = done_moves < scramble_nums
        while np.any(lt):
            idxs = np.where(lt)[0]
            idxs = np.random.choice(idxs, int(max(len(idxs) / 12, 1)))
            move = randrange(12)
            di = torch.from_numpy(idxs).to(st.device)
            st[di] = ops.next_state(self.env_id, st[di].contiguous(), move)
            done_moves[idxs] = done_moves[idxs] + 1
            lt[idxs] = done_moves[idxs] < scramble_nums[idxs]
        return self.unpack(st.cpu().numpy()), scramble_nums.tolist()

    def expand(self, states: List[State]) -> Tuple[List[List[State]], List[np.ndarray]]:
        n = len(states)
        ch = self._expand_np(self.pack(states)).astype(self.dtype, copy=False)      # [N,12,54]
        children = [[Cube3State(ch[i, a]) for a in range(12)] for i in range(n)]
        tc = np.ones([n, 12])
        return children, [tc[i] for i in range(n)]
