"""Lights Out (7x7) behind the reference's Environment API (environments/lights_out.py:9-166), GPU-backed.

State = dim*dim cells of 0/1 (uint8), goal = all off; move m presses cell m: it and its in-board neighbours toggle
(move_matrix rows [m, m+dim, m-dim, m+1, m-1], an out-of-board neighbour replaced by m itself, lights_out.py:31-42).
A press is its own inverse.  The CUDA path works on the 49-bit form of the board (csrc/lightsout_kernels.cu).
SURVEY 8(f) rank 4: same search engine, same ABI, new kernel shape (49 children of 49 bytes per parent).
"""
from __future__ import annotations

from typing import List, Tuple, Union

import numpy as np
from torch import nn

from .. import ops
from .._lib import ENV_IDS
from ..utils.pytorch_models import ResnetModel
from ._packed import PackedEnvMixin
from .environment_abstract import Environment, State


class LOState(State):
    __slots__ = ["tiles", "hash"]

    def __init__(self, tiles: np.ndarray):
        self.tiles: np.ndarray = tiles
        self.hash = None

    def __hash__(self):
        h = getattr(self, "hash", None)
        if h is None:
            h = hash(np.asarray(self.tiles).tobytes())
            self.hash = h
        return h

    def __eq__(self, other):
        return np.array_equal(self.tiles, other.tiles)


LOState.__module__ = "environments.lights_out"


class LightsOut(PackedEnvMixin, Environment):
    _state_cls = LOState
    _attr = "tiles"

    def __init__(self, dim: int):
        super().__init__()
        if dim != 7:
            raise ValueError("the CUDA path is compiled for the 7x7 board (lightsout7), got %d" % dim)
        self.dtype = np.uint8
        self.dim = dim
        self.num_tiles = dim * dim
        self.state_dim = self.num_tiles
        self.env_id = ENV_IDS["lightsout%d" % dim]
        cell = np.arange(self.num_tiles)
        x, y = cell // dim, cell % dim
        self.move_matrix = np.stack([cell, np.where(x < dim - 1, cell + dim, cell), np.where(x > 0, cell - dim, cell),
                                     np.where(y < dim - 1, cell + 1, cell), np.where(y > 0, cell - 1, cell)], axis=1).astype(np.int64)

    def next_state(self, states: List[LOState], action: int) -> Tuple[List[LOState], List[float]]:
        nxt, tcs = self._next_state_np(self.pack(states), action)
        return self.unpack(nxt), tcs

    def prev_state(self, states: List[LOState], action: int) -> List[LOState]:
        return self.next_state(states, action)[0]

    def generate_goal_states(self, num_states: int, np_format: bool = False) -> Union[List[LOState], np.ndarray]:
        if np_format:
            return np.zeros((num_states, self.num_tiles), dtype=self.dtype)
        return [LOState(np.zeros(self.num_tiles, dtype=self.dtype)) for _ in range(num_states)]

    def is_solved(self, states: List[LOState]) -> np.ndarray:
        return self._is_solved_np(self.pack(states))

    def state_to_nnet_input(self, states: List[LOState]) -> List[np.ndarray]:
        x = ops.nnet_input(self.env_id, self.to_device(self.pack(states))).cpu().numpy()
        return [x.astype(self.dtype, copy=False)]

    def get_num_moves(self) -> int:
        return self.num_tiles

    def get_nnet_model(self) -> nn.Module:
        return ResnetModel(self.num_tiles, 6, 5000, 1000, 4, 1, True)

    def generate_states(self, num_states: int, backwards_range: Tuple[int, int]) -> Tuple[List[LOState], List[int]]:
        """lights_out.py:86-119: a pre-drawn move table, all unfinished states advance together (reference RNG sequence);
        the per-state moves of one round are applied on the GPU, one launch per distinct move."""
        assert num_states > 0 and backwards_range[0] >= 0
        import torch
        depths = list(range(backwards_range[0], backwards_range[1] + 1))
        st = self.to_device(self.generate_goal_states(num_states, np_format=True))
        scramble_nums = np.random.choice(depths, num_states)
        done_moves = np.zeros(num_states)
        moves = np.random.choice(self.num_tiles, size=(num_states, max(depths)))
        k = 0
        lt = done_moves < scramble_nums
        while np.any(lt):
            idxs = np.where(lt)[0]
            mv = moves[idxs, k]
            for a in np.unique(mv):
                di = torch.from_numpy(idxs[mv == a]).to(st.device)
                st[di] = ops.next_state(self.env_id, st[di].contiguous(), int(a))
            done_moves[idxs] = done_moves[idxs] + 1
            lt[idxs] = done_moves[idxs] < scramble_nums[idxs]
            k += 1
        return self.unpack(st.cpu().numpy()), scramble_nums.tolist()

    def expand(self, states: List[State]) -> Tuple[List[List[State]], List[np.ndarray]]:
        n, a = len(states), self.num_tiles
        ch = self._expand_np(self.pack(states)).astype(self.dtype, copy=False)
        children = [[LOState(ch[i, m]) for m in range(a)] for i in range(n)]
        tc = np.ones([n, a])
        return children, [tc[i] for i in range(n)]
