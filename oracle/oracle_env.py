"""numpy restatement of the reference's environment step (cube3 / n-puzzle).  TEST INFRASTRUCTURE ONLY.

Pinned against tests/golden/* (generated from the unmodified reference by tests/golden/make_golden.py):
tables, 10k seeded scrambles x 12 moves (config 1), the 21,349 / 26,011 / 127,835 (s,a,s') triples of the
shipped results, and the optimal solutions shipped with data/*/test.  `state_hash64` has no reference
counterpart (reference hashes are CPython's salted bytes-hash / boost::hash_range, both unobservable):
it is the definition the CUDA path must reproduce bit-exactly -- hash values: parity unpinned by design.
"""
from __future__ import annotations

import json
import os
import random
from typing import List, Tuple

import numpy as np

_GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class OracleCube3:
    """environments/cube3.py:27-171 and cpp/environments.cpp:222-256."""
    name = "cube3"
    state_dim = 54
    num_moves = 12

    def __init__(self):
        t = json.load(open(os.path.join(_GOLD, "cube3_tables.json")))
        self.moves: List[str] = t["moves"]
        self.moves_rev: List[str] = t["moves_rev"]
        self.idxs_new = [np.array(x) for x in t["idxs_new"]]   # cube3.py:183-256 / environments.h:91-104
        self.idxs_old = [np.array(x) for x in t["idxs_old"]]   # environments.h:77-90
        self.goal = np.arange(54, dtype=np.uint8)               # cube3.py:37
        self.rev_action = [self.moves_rev.index(m) for m in self.moves]  # cube3.py:56-60

    def move(self, states: np.ndarray, action: int) -> np.ndarray:
        """cube3.py:163-171 `_move_np`: copy, then next[:, new] = cur[:, old]."""
        nxt = states.copy()
        nxt[:, self.idxs_new[action]] = states[:, self.idxs_old[action]]
        return nxt

    def prev(self, states: np.ndarray, action: int) -> np.ndarray:
        return self.move(states, self.rev_action[action])

    def is_solved(self, states: np.ndarray) -> np.ndarray:
        """cube3.py:71-75: sticker identity (state[i] == i), not colour equivalence."""
        return np.all(states == self.goal[None, :], axis=1)

    def nnet_input(self, states: np.ndarray) -> np.ndarray:
        """cube3.py:77-85: (colors / 9).astype(uint8) -> colour id 0..5."""
        return (states / 9).astype(np.uint8)

    def expand(self, states: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """cube3.py:129-161: children[N,12,54] parent-major / move-minor, transition cost 1.0 each."""
        ch = np.stack([self.move(states, a) for a in range(12)], axis=1)
        return ch, np.ones((states.shape[0], 12), dtype=np.float64)

    def generate_states(self, n: int, back: Tuple[int, int]) -> Tuple[np.ndarray, np.ndarray]:
        """cube3.py:96-127: same numpy / `random` call sequence, so fixed seeds reproduce the reference."""
        scrambs = list(range(back[0], back[1] + 1))
        st = np.repeat(self.goal[None, :].copy(), n, axis=0)
        scr = np.random.choice(scrambs, n)
        nb = np.zeros(n)
        lt = nb < scr
        while np.any(lt):
            idxs = np.where(lt)[0]
            sub = int(max(len(idxs) / 12, 1))
            idxs = np.random.choice(idxs, sub)
            mv = random.randrange(12)
            st[idxs] = self.move(st[idxs], mv)
            nb[idxs] = nb[idxs] + 1
            lt[idxs] = nb[idxs] < scr[idxs]
        return st, scr


class OracleNPuzzle:
    """environments/n_puzzle.py:27-231 and cpp/environments.cpp:4-126."""
    num_moves = 4

    def __init__(self, dim: int):
        self.dim = dim
        self.state_dim = dim * dim
        self.name = "puzzle%d" % (dim * dim - 1)
        t = json.load(open(os.path.join(_GOLD, "puzzle_tables.json")))[str(dim)]
        self.moves, self.moves_rev = t["moves"], t["moves_rev"]
        self.swap = np.array(t["swap_zero_idxs"], dtype=np.int64)   # n_puzzle.py:174-214
        self.goal = np.array(t["goal"], dtype=np.uint8)             # n_puzzle.py:41
        self.rev_action = [self.moves_rev.index(m) for m in self.moves]

    @staticmethod
    def build_swap_table(dim: int) -> np.ndarray:
        """Independent restatement of n_puzzle.py:174-214 / environments.cpp:4-46 (checked against golden)."""
        t = np.zeros((dim * dim, 4), dtype=np.int64)
        for i in range(dim):
            for j in range(dim):
                z = i * dim + j
                t[z, 0] = (i + 1) * dim + j if i < dim - 1 else z   # U: blank swaps with the tile below
                t[z, 1] = (i - 1) * dim + j if i > 0 else z         # D
                t[z, 2] = i * dim + j + 1 if j < dim - 1 else z     # L
                t[z, 3] = i * dim + j - 1 if j > 0 else z           # R
        return t

    def move(self, states: np.ndarray, action: int) -> np.ndarray:
        """n_puzzle.py:216-231: next[z] = cur[swap]; next[swap] = 0 (illegal move => swap == z => no-op)."""
        nxt = states.copy()
        rows = np.arange(states.shape[0])
        z = np.where(states == 0)[1]
        s = self.swap[z, action]
        nxt[rows, z] = states[rows, s]
        nxt[rows, s] = 0
        return nxt

    def prev(self, states: np.ndarray, action: int) -> np.ndarray:
        return self.move(states, self.rev_action[action])

    def is_solved(self, states: np.ndarray) -> np.ndarray:
        """n_puzzle.py:78-82."""
        return np.all(states == self.goal[None, :], axis=1)

    def nnet_input(self, states: np.ndarray) -> np.ndarray:
        """n_puzzle.py:84-89: raw tiles."""
        return states.astype(np.uint8)

    def expand(self, states: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        ch = np.stack([self.move(states, a) for a in range(4)], axis=1)
        return ch, np.ones((states.shape[0], 4), dtype=np.float64)

    def generate_states(self, n: int, back: Tuple[int, int]) -> Tuple[np.ndarray, np.ndarray]:
        """n_puzzle.py:100-134 (same RNG call sequence)."""
        scrambs = list(range(back[0], back[1] + 1))
        st = np.repeat(self.goal[None, :].copy(), n, axis=0)
        scr = np.random.choice(scrambs, n)
        nb = np.zeros(n)
        while np.max(nb < scr):
            idxs = np.where(nb < scr)[0]
            sub = int(max(len(idxs) / 4, 1))
            idxs = np.random.choice(idxs, sub)
            mv = random.randrange(4)
            st[idxs] = self.move(st[idxs], mv)
            nb[idxs] = nb[idxs] + 1
        return st, scr


class OracleLightsOut:
    """environments/lights_out.py:26-166 and cpp/environments.cpp:133-208."""

    def __init__(self, dim: int = 7):
        self.dim = dim
        self.state_dim = dim * dim
        self.num_moves = dim * dim
        self.name = "lightsout%d" % dim
        self.move_matrix = self.build_move_matrix(dim)                   # lights_out.py:31-42
        self.goal = np.zeros(self.state_dim, dtype=np.uint8)             # lights_out.py:56-64
        self.rev_action = list(range(self.num_moves))                    # a press undoes itself (lights_out.py:53-54)

    @staticmethod
    def build_move_matrix(dim: int) -> np.ndarray:
        """[move, move+dim, move-dim, move+1, move-1] with an out-of-board neighbour replaced by `move` itself."""
        mm = np.zeros((dim * dim, 5), dtype=np.int64)
        for move in range(dim * dim):
            x, y = move // dim, move % dim
            mm[move] = [move, move + dim if x < dim - 1 else move, move - dim if x > 0 else move,
                        move + 1 if y < dim - 1 else move, move - 1 if y > 0 else move]
        return mm

    def move_many(self, states: np.ndarray, actions) -> np.ndarray:
        """lights_out.py:156-166 `_move_np`: every listed cell becomes (old + 1) % 2 of the ORIGINAL state, so a cell listed
        twice (board edge) still toggles once -- same as LightsOut::getNextState (environments.cpp:171-183)."""
        nxt = states.copy()
        rows = np.arange(states.shape[0])[:, None]
        mm = self.move_matrix[np.asarray(actions)]
        nxt[rows, mm] = (states[rows, mm] + 1) % 2
        return nxt

    def move(self, states: np.ndarray, action: int) -> np.ndarray:
        return self.move_many(states, [action] * states.shape[0])

    def prev(self, states: np.ndarray, action: int) -> np.ndarray:
        return self.move(states, action)

    def is_solved(self, states: np.ndarray) -> np.ndarray:
        return np.all(states == 0, axis=1)                               # lights_out.py:66-69

    def nnet_input(self, states: np.ndarray) -> np.ndarray:
        return states.astype(np.uint8)                                   # lights_out.py:71-76

    def expand(self, states: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        ch = np.stack([self.move(states, a) for a in range(self.num_moves)], axis=1)
        return ch, np.ones((states.shape[0], self.num_moves), dtype=np.float64)

    def generate_states(self, n: int, back: Tuple[int, int]) -> Tuple[np.ndarray, np.ndarray]:
        """lights_out.py:86-119: one pre-drawn move table, states advance in lock step (same RNG call sequence)."""
        scrambs = list(range(back[0], back[1] + 1))
        st = np.zeros((n, self.state_dim), dtype=np.uint8)
        scr = np.random.choice(scrambs, n)
        nb = np.zeros(n)
        moves = np.random.choice(self.num_moves, size=(n, max(scrambs)))
        k = 0
        lt = nb < scr
        while np.any(lt):
            idxs = np.where(lt)[0]
            st[idxs] = self.move_many(st[idxs], list(moves[idxs, k]))
            nb[idxs] = nb[idxs] + 1
            lt[idxs] = nb[idxs] < scr[idxs]
            k += 1
        return st, scr


class OracleCube4:
    """cpp/environments.cpp:262-370 (the reference has no Python Cube4).  96 sticker ids, 24 quarter turns
    (12 outer layers, then 12 inner slices); solved = every face shows one colour (state[i] / 16), environments.cpp:356-366."""
    name = "cube4"
    state_dim = 96
    num_moves = 24

    def __init__(self):
        t = json.load(open(os.path.join(_GOLD, "cube4_tables.json")))
        self.perm = np.array(t["perm"], dtype=np.int64)          # child[j] = parent[perm[a][j]] == newState[new] = state[old]
        self.goal = np.arange(96, dtype=np.uint8)
        self.rev_action = [a ^ 1 for a in range(24)]             # the opposite quarter turn of the same layer

    def move(self, states: np.ndarray, action: int) -> np.ndarray:
        """Cube4::getNextState (environments.cpp:327-341) in gather form."""
        return np.ascontiguousarray(states[:, self.perm[action]])     # (fancy indexing alone hands back a transposed layout)

    def prev(self, states: np.ndarray, action: int) -> np.ndarray:
        return self.move(states, self.rev_action[action])

    def is_solved(self, states: np.ndarray) -> np.ndarray:
        """Cube4::isSolved (environments.cpp:356-366): colour = sticker id / 16, all 16 stickers of a face equal."""
        col = (states // 16).reshape(states.shape[0], 6, 16)
        return np.all(col == col[:, :, :1], axis=(1, 2))

    def nnet_input(self, states: np.ndarray) -> np.ndarray:
        """By analogy with cube3.py:77-85 (sticker id -> colour id); the reference ships no cube4 network."""
        return (states // 16).astype(np.uint8)

    def expand(self, states: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        ch = np.ascontiguousarray(np.stack([self.move(states, a) for a in range(24)], axis=1))
        return ch, np.ones((states.shape[0], 24), dtype=np.float64)

    def generate_states(self, n: int, back: Tuple[int, int]) -> Tuple[np.ndarray, np.ndarray]:
        """The cube3 scrambler (cube3.py:96-127) with 24 moves -- same numpy / `random` call sequence."""
        scrambs = list(range(back[0], back[1] + 1))
        st = np.repeat(self.goal[None, :].copy(), n, axis=0)
        scr = np.random.choice(scrambs, n)
        nb = np.zeros(n)
        lt = nb < scr
        while np.any(lt):
            idxs = np.where(lt)[0]
            sub = int(max(len(idxs) / 24, 1))
            idxs = np.random.choice(idxs, sub)
            mv = random.randrange(24)
            st[idxs] = self.move(st[idxs], mv)
            nb[idxs] = nb[idxs] + 1
            lt[idxs] = nb[idxs] < scr[idxs]
        return st, scr


def get_oracle_env(name: str):
    """utils/env_utils.py:6-28 (cube3, puzzle(\\d+), lightsout(\\d+)) + cube4 (parallel_weighted_astar.cpp:386)."""
    name = name.lower()
    if name == "cube3":
        return OracleCube3()
    if name == "cube4":
        return OracleCube4()
    if name.startswith("lightsout"):
        return OracleLightsOut(int(name[9:]))
    if name.startswith("puzzle"):
        return OracleNPuzzle(int(round((int(name[6:]) + 1) ** 0.5)))
    raise ValueError("No known environment %s" % name)


# ---------------------------------------------------------------------------------------------------
# State hash (project-defined; see module docstring).  NH-style pair-product universal hash over the
# state's bytes zero-padded to a multiple of 8, little-endian u32 words, then murmur3's fmix64.
# ---------------------------------------------------------------------------------------------------
HASH_SEED = np.uint64(0x9E3779B97F4A7C15)


def _splitmix64_keys(n: int) -> np.ndarray:
    x = 0x0DCB2000DCB200
    out = []
    for _ in range(n):
        x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        z = z ^ (z >> 31)
        out.append((z >> 32) | 1)
    return np.array(out, dtype=np.uint64)


HASH_KEYS = _splitmix64_keys(24)   # 32-bit odd keys, enough for states up to 96 bytes (cube4)


def hash_words(state_dim: int) -> int:
    return 2 * ((state_dim + 7) // 8)


def state_hash64(states: np.ndarray) -> np.ndarray:
    n, s = states.shape
    w = hash_words(s)
    buf = np.zeros((n, 4 * w), dtype=np.uint8)
    buf[:, :s] = states
    words = buf.view("<u4").astype(np.uint64)                       # [n, w]
    m = np.uint64(0xFFFFFFFF)
    acc = np.full(n, HASH_SEED, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for i in range(0, w, 2):
            a = (words[:, i] + HASH_KEYS[i]) & m
            b = (words[:, i + 1] + HASH_KEYS[i + 1]) & m
            acc = acc + a * b
        acc ^= acc >> np.uint64(33)
        acc *= np.uint64(0xFF51AFD7ED558CCD)
        acc ^= acc >> np.uint64(33)
        acc *= np.uint64(0xC4CEB9FE1A85EC53)
        acc ^= acc >> np.uint64(33)
    acc[acc == 0] = np.uint64(1)                                     # 0 is the closed table's EMPTY key
    return acc
