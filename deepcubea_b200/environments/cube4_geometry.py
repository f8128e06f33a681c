"""Geometric construction of the 4x4x4 cube's 24 move permutations.

The reference has Cube4 only in C++ and ships its moves as index literals (cpp/environments.cpp:262-317:
`rotateIdxs_old` / `rotateIdxs_new`, applied as newState[new] = state[old], :327-341).  Here the 96 stickers are
embedded in 3-D with the SAME face frames as the 3x3x3 cube (cube3_geometry.FACE_FRAMES; sticker index =
16*face + 4*i + j, faces U,D,L,R,B,F) and a move is a quarter turn of one layer about its face's outward normal:
depth 0 = the outer layer (16 face stickers + 16 on the adjacent strips), depth 1 = the inner slice next to it
(16 stickers).  tests/test_host_logic.py pins the result to tests/golden/cube4_tables.json, which
tests/golden/make_golden_cube4.py dumps from the compiled, unmodified reference class.

perm[a][j] = index of the parent sticker that lands on position j, i.e. child[j] = parent[perm[a][j]].
Move order (environments.cpp:289): U0-1 U0+1 D0-1 D0+1 L0.. R0.. B0.. F0.., then U1-1 U1+1 ... F1+1.
"""
from __future__ import annotations

from typing import List

import numpy as np

from .cube3_geometry import FACE_FRAMES, FACES, _quarter_turn

N = 4
NUM_STICKERS = 6 * N * N
MOVES: List[str] = ["%s%d%+d" % (f, depth, n) for depth in (0, 1) for f in FACES for n in (-1, 1)]


def sticker_positions() -> np.ndarray:
    """[96,3] doubled coordinates: cubie centres at -3,-1,1,3 along a face, sticker plane at +-4."""
    pos = np.zeros((NUM_STICKERS, 3), dtype=np.int64)
    for f, (n, u, v) in enumerate(FACE_FRAMES):
        n, u, v = np.array(n), np.array(u), np.array(v)
        for i in range(N):
            for j in range(N):
                pos[N * N * f + N * i + j] = N * n + (2 * i - (N - 1)) * u + (2 * j - (N - 1)) * v
    return pos


def move_permutations() -> np.ndarray:
    """perm[24][96] (int64), child[j] = parent[perm[a][j]]."""
    pos = sticker_positions()
    index_at = {tuple(p): k for k, p in enumerate(pos)}
    perm = np.tile(np.arange(NUM_STICKERS), (2 * 12, 1))
    for depth in (0, 1):
        for f, (n, _, _) in enumerate(FACE_FRAMES):
            n = np.array(n)
            height = pos @ n
            on_layer = height >= N - 1 if depth == 0 else height == N - 1 - 2 * depth
            for k, sign in enumerate((-1, 1)):
                rot = _quarter_turn(n, -sign)       # "+1" is clockwise seen from outside, as for cube3
                for src in np.nonzero(on_layer)[0]:
                    perm[12 * depth + 2 * f + k, index_at[tuple(rot @ pos[src])]] = src
    return perm


def inverse_actions() -> List[int]:
    """action -> the action that undoes it (the opposite quarter turn of the same layer)."""
    return [a ^ 1 for a in range(24)]
