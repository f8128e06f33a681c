"""(n^2-1)-puzzle behind the reference's Environment API (environments/n_puzzle.py:10-231), GPU-backed.

State = n*n tiles (0 = blank), goal = [1..n*n-1, 0]; moves U,D,L,R swap the blank with the tile at
(i+1,j),(i-1,j),(i,j+1),(i,j-1); an illegal move is a no-op that still costs 1.  dim 4..7 (puzzle15..48).
"""
from __future__ import annotations

from random import randrange
from typing import List, Tuple, Union

import numpy as np
import torch.nn as nn

from .. import ops
from .._lib import ENV_IDS
from ..utils.pytorch_models import ResnetModel
from ._packed import PackedEnvMixin
from .environment_abstract import Environment, State


class NPuzzleState(State):
    __slots__ = ["tiles", "hash"]

    def __init__(self, tiles: np.ndarray):
        self.tiles: np.ndarray = tiles
        self.hash = None

    def __hash__(self):
        h = getattr(self, "hash", None)     # puzzle15/24 pickles predate the slot
        if h is None:
            h = hash(np.asarray(self.tiles).tobytes())
            self.hash = h
        return h

    def __eq__(self, other):
        return np.array_equal(self.tiles, other.tiles)


NPuzzleState.__module__ = "environments.n_puzzle"


class NPuzzle(PackedEnvMixin, Environment):
    moves: List[str] = ["U", "D", "L", "R"]
    moves_rev: List[str] = ["D", "U", "R", "L"]
    _state_cls = NPuzzleState
    _attr = "tiles"

    def __init__(self, dim: int):
        super().__init__()
        if dim not in (4, 5, 6, 7):
            raise ValueError("the CUDA path is compiled for dim 4..7 (puzzle15/24/35/48), got %d" % dim)
        self.dim = dim
        self.state_dim = dim * dim
        self.env_id = ENV_IDS["puzzle%d" % (dim * dim - 1)]
        self.dtype = np.uint8
        self.goal_tiles: np.ndarray = np.concatenate((np.arange(1, dim * dim), [0])).astype(self.dtype)
        self.swap_zero_idxs: np.ndarray = self._get_swap_zero_idxs(dim)

    def next_state(self, states: List[NPuzzleState], action: int) -> Tuple[List[NPuzzleState], List[float]]:
        nxt, tcs = self._next_state_np(self.pack(states), action)
        return self.unpack(nxt), tcs

    def prev_state(self, states: List[NPuzzleState], action: int) -> List[NPuzzleState]:
        return self.next_state(states, self.moves_rev.index(self.moves[action]))[0]

    def generate_goal_states(self, num_states: int, np_format: bool = False) -> Union[List[NPuzzleState], np.ndarray]:
        if np_format:
            return np.repeat(self.goal_tiles[None, :].copy(), num_states, axis=0)
        return [NPuzzleState(self.goal_tiles.copy()) for _ in range(num_states)]

    def is_solved(self, states: List[NPuzzleState]) -> np.ndarray:
        return self._is_solved_np(self.pack(states))

    def state_to_nnet_input(self, states: List[NPuzzleState]) -> List[np.ndarray]:
        x = ops.nnet_input(self.env_id, self.to_device(self.pack(states))).cpu().numpy()
        return [x.astype(self.dtype, copy=False)]

    def get_num_moves(self) -> int:
        return 4

    def get_nnet_model(self) -> nn.Module:
        return ResnetModel(self.state_dim, self.dim ** 2, 5000, 1000, 4, 1, True)

    def generate_states(self, num_states: int, backwards_range: Tuple[int, int]) -> Tuple[List[NPuzzleState], List[int]]:
        """n_puzzle.py:100-134, state array on the GPU, reference RNG call sequence."""
        assert num_states > 0 and backwards_range[0] >= 0
        import torch
        depths = list(range(backwards_range[0], backwards_range[1] + 1))
        st = self.to_device(self.generate_goal_states(num_states, np_format=True))
        scramble_nums = np.random.choice(depths, num_states)
        done_moves = np.zeros(num_states)
        while np.max(done_moves < scramble_nums):
            idxs = np.where(done_moves < scramble_nums)[0]
            idxs = np.random.choice(idxs, int(max(len(idxs) / 4, 1)))
            move = randrange(4)
            di = torch.from_numpy(idxs).to(st.device)
            st[di] = ops.next_state(self.env_id, st[di].contiguous(), move)
            done_moves[idxs] = done_moves[idxs] + 1
        return self.unpack(st.cpu().numpy()), scramble_nums.tolist()

    def expand(self, states: List[State]) -> Tuple[List[List[State]], List[np.ndarray]]:
        n = len(states)
        ch = self._expand_np(self.pack(states)).astype(self.dtype, copy=False)
        children = [[NPuzzleState(ch[i, a]) for a in range(4)] for i in range(n)]
        tc = np.ones([n, 4])
        return children, [tc[i] for i in range(n)]

    def _get_swap_zero_idxs(self, n: int) -> np.ndarray:
        """Blank-swap table [n*n,4] (n_puzzle.py:174-214): target index, or the blank's own index if illegal."""
        z = np.arange(n * n)
        row, col = z // n, z % n
        table = np.stack([np.where(row < n - 1, z + n, z), np.where(row > 0, z - n, z),
                          np.where(col < n - 1, z + 1, z), np.where(col > 0, z - 1, z)], axis=1)
        return table.astype(self.dtype)
