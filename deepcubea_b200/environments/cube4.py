"""4x4x4 cube behind the Environment API, GPU-backed.

The reference has this environment only as the C++ class `Cube4` (cpp/environments.cpp:262-370), selected by
`parallel_weighted_astar.cpp:386`; there is no environments/cube4.py and no trained network.  This class gives it the
same plugin surface as Cube3 so the CUDA search (`astar.py --language cuda --env cube4`) and the Environment methods work
on it.  State = 96 sticker ids (uint8); 24 moves = 12 outer-layer quarter turns then 12 inner-slice quarter turns
(cube4_geometry.MOVES, the order of environments.cpp:289); **solved = every face shows one colour (id // 16)**
(Cube4::isSolved, environments.cpp:356-366) -- not sticker identity.  The network input follows cube3.py:77-85 with 16
stickers per face; `get_nnet_model` mirrors cube3's architecture for a 96-sticker input (random-init only).
"""
from __future__ import annotations

from random import randrange
from typing import List, Tuple, Union

import numpy as np
from torch import nn

from .. import ops
from .._lib import ENV_IDS
from ..utils.pytorch_models import ResnetModel
from . import cube4_geometry
from ._packed import PackedEnvMixin
from .environment_abstract import Environment, State


class Cube4State(State):
    __slots__ = ["colors", "hash"]

    def __init__(self, colors: np.ndarray):
        self.colors: np.ndarray = colors
        self.hash = None

    def __hash__(self):
        if self.hash is None:
            self.hash = hash(np.asarray(self.colors).tobytes())
        return self.hash

    def __eq__(self, other):
        return np.array_equal(self.colors, other.colors)


class Cube4(PackedEnvMixin, Environment):
    moves: List[str] = cube4_geometry.MOVES
    env_id = ENV_IDS["cube4"]
    state_dim = 96
    _state_cls = Cube4State
    _attr = "colors"

    def __init__(self):
        super().__init__()
        self.dtype = np.uint8
        self.cube_len = 4
        self.goal_colors: np.ndarray = np.arange(0, 96, 1, dtype=self.dtype)
        self._rev_action = cube4_geometry.inverse_actions()

    def next_state(self, states: List[Cube4State], action: int) -> Tuple[List[Cube4State], List[float]]:
        nxt, tcs = self._next_state_np(self.pack(states), action)
        return self.unpack(nxt), tcs

    def prev_state(self, states: List[Cube4State], action: int) -> List[Cube4State]:
        return self.next_state(states, self._rev_action[action])[0]

    def generate_goal_states(self, num_states: int, np_format: bool = False) -> Union[List[Cube4State], np.ndarray]:
        if np_format:
            return np.repeat(self.goal_colors[None, :].copy(), num_states, axis=0)
        return [Cube4State(self.goal_colors.copy()) for _ in range(num_states)]

    def is_solved(self, states: List[Cube4State]) -> np.ndarray:
        return self._is_solved_np(self.pack(states))

    def state_to_nnet_input(self, states: List[Cube4State]) -> List[np.ndarray]:
        x = ops.nnet_input(self.env_id, self.to_device(self.pack(states))).cpu().numpy()
        return [x.astype(self.dtype, copy=False)]

    def get_num_moves(self) -> int:
        return len(self.moves)

    def get_nnet_model(self) -> nn.Module:
        return ResnetModel(96, 6, 5000, 1000, 4, 1, True)

    def generate_states(self, num_states: int, backwards_range: Tuple[int, int]) -> Tuple[List[Cube4State], List[int]]:
        """The cube3 scrambler (cube3.py:96-127) over 24 moves, state array resident on the GPU."""
        assert num_states > 0 and backwards_range[0] >= 0
        import torch
        depths = list(range(backwards_range[0], backwards_range[1] + 1))
        st = self.to_device(self.generate_goal_states(num_states, np_format=True))
        scramble_nums = np.random.choice(depths, num_states)
        done_moves = np.zeros(num_states)
        lt = done_moves < scramble_nums
        while np.any(lt):
            idxs = np.where(lt)[0]
            idxs = np.random.choice(idxs, int(max(len(idxs) / 24, 1)))
            move = randrange(24)
            di = torch.from_numpy(idxs).to(st.device)
            st[di] = ops.next_state(self.env_id, st[di].contiguous(), move)
            done_moves[idxs] = done_moves[idxs] + 1
            lt[idxs] = done_moves[idxs] < scramble_nums[idxs]
        return self.unpack(st.cpu().numpy()), scramble_nums.tolist()

    def expand(self, states: List[State]) -> Tuple[List[List[State]], List[np.ndarray]]:
        n = len(states)
        ch = self._expand_np(self.pack(states)).astype(self.dtype, copy=False)      # [N,24,96]
        children = [[Cube4State(ch[i, a]) for a in range(24)] for i in range(n)]
        tc = np.ones([n, 24])
        return children, [tc[i] for i in range(n)]
