"""Geometric construction of the 3x3x3 cube's move permutations.

The reference builds its tables from hand-written per-face index lists (environments/cube3.py:183-256) or
ships them as literals (cpp/environments.h:75-105).  Here the 54 stickers are embedded in 3-D and a move is
what it physically is -- a quarter turn of one face layer about that face's outward normal -- so the table
falls out of integer rotation matrices.  Sticker numbering is the reference's: index = 9*face + 3*i + j with
faces U,D,L,R,B,F = 0..5 (cube3.py:28, 220).  `FACE_FRAMES` gives, per face, the outward normal and the
directions in which i and j grow; it was identified by tools/derive_cube3_geometry.py as the embedding that
reproduces the reference's permutations, and tests/test_tables.py pins the result to tests/golden.

perm[a][j] = index of the parent sticker that lands on position j after move a, i.e.
child[j] = parent[perm[a][j]]  (the gather form of `next[:, idxs_new] = cur[:, idxs_old]`, cube3.py:167).
Move order: U-1 U1 D-1 D1 L-1 L1 R-1 R1 B-1 B1 F-1 F1 (cube3.py:28).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

FACES = "UDLRBF"
MOVES: List[str] = ["%s%i" % (f, n) for f in FACES for n in (-1, 1)]
# the inverse of move "X-1" is "X1" and vice versa
MOVES_REV: List[str] = ["%s%i" % (f, n) for f in FACES for n in (1, -1)]

# face -> (outward normal, direction of increasing i, direction of increasing j)
FACE_FRAMES: Tuple[Tuple[Tuple[int, int, int], ...], ...] = (
    ((0, 0, 1), (1, 0, 0), (0, 1, 0)),     # U
    ((0, 0, -1), (1, 0, 0), (0, -1, 0)),   # D
    ((-1, 0, 0), (0, -1, 0), (0, 0, 1)),   # L
    ((1, 0, 0), (0, 1, 0), (0, 0, 1)),     # R
    ((0, 1, 0), (-1, 0, 0), (0, 0, 1)),    # B
    ((0, -1, 0), (1, 0, 0), (0, 0, 1)),    # F
)


def _quarter_turn(axis: np.ndarray, sign: int) -> np.ndarray:
    """Integer rotation matrix for sign*90 degrees about a unit axis (Rodrigues with cos=0, sin=sign)."""
    x, y, z = axis
    cross = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    return np.outer(axis, axis) + sign * cross


def sticker_positions() -> np.ndarray:
    """[54,3] doubled coordinates: cubie centres at -2/0/2 along the face, sticker plane at +-3."""
    pos = np.zeros((54, 3), dtype=np.int64)
    for f, (n, u, v) in enumerate(FACE_FRAMES):
        n, u, v = np.array(n), np.array(u), np.array(v)
        for i in range(3):
            for j in range(3):
                pos[9 * f + 3 * i + j] = 3 * n + 2 * (i - 1) * u + 2 * (j - 1) * v
    return pos


def move_permutations() -> np.ndarray:
    """perm[12][54] (int64), child[j] = parent[perm[a][j]]."""
    pos = sticker_positions()
    index_at = {tuple(p): k for k, p in enumerate(pos)}
    perm = np.tile(np.arange(54), (12, 1))
    for f, (n, _, _) in enumerate(FACE_FRAMES):
        n = np.array(n)
        on_layer = pos @ n >= 2                       # the 9 face stickers + the 12 adjacent strip stickers
        for k, sign in enumerate((-1, 1)):
            # the reference's "+1" turn is clockwise seen from outside = -90 degrees about the outward normal
            rot = _quarter_turn(n, -sign)
            for src in np.nonzero(on_layer)[0]:
                dst = index_at[tuple(rot @ pos[src])]
                perm[2 * f + k, dst] = src
    return perm


def inverse_actions() -> List[int]:
    """action -> action that undoes it (Cube3.prev_state, cube3.py:56-60)."""
    return [MOVES_REV.index(m) for m in MOVES]
