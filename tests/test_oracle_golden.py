"""CPU tier: pin the oracle against fixtures produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
import ctypes
import hashlib
import json
import random

import numpy as np
import pytest

from oracle import oracle_env as O


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_cube3_tables_are_permutations(golden_dir):
    t = json.load(open(golden_dir + "/cube3_tables.json"))
    perm = np.array(t["perm"])
    assert perm.shape == (12, 54)
    for a in range(12):
        assert sorted(perm[a]) == list(range(54))
        assert (perm[a] != np.arange(54)).sum() == 20                      # 20 stickers move, 34 stay
        inv = perm[a ^ 1]
        assert np.array_equal(perm[a][inv], np.arange(54))                 # (2k, 2k+1) are inverse pairs
        p4 = np.arange(54)
        for _ in range(4):
            p4 = p4[perm[a]]
        assert np.array_equal(p4, np.arange(54))                           # quarter turn ^4 = identity


def test_cube3_config1_reproduces_reference(golden_dir):
    g = np.load(golden_dir + "/cube3_cfg1.npz")
    env = O.OracleCube3()
    np.random.seed(0); random.seed(0)
    st, depths = env.generate_states(10000, (0, 26))
    assert np.array_equal(st, g["parents"]) and np.array_equal(depths, g["depths"])
    ch, tc = env.expand(st)
    assert _sha(ch) == str(g["children_sha256"])
    assert np.array_equal(ch[:256], g["children_head"])
    assert (tc == 1.0).all()
    solved = env.is_solved(ch.reshape(-1, 54)).reshape(10000, 12)
    assert np.array_equal(np.packbits(solved), g["solved"]) and solved.sum() == int(g["n_solved"])
    assert _sha(env.nnet_input(st)) == str(g["nnet_in_sha256"])


@pytest.mark.parametrize("name,dim", [("puzzle15", 4), ("puzzle48", 7)])
def test_puzzle_config1_reproduces_reference(golden_dir, name, dim):
    g = np.load(golden_dir + "/%s_cfg1.npz" % name)
    env = O.OracleNPuzzle(dim)
    np.random.seed(1); random.seed(1)
    st, depths = env.generate_states(2000, (0, 60))
    assert np.array_equal(st, g["parents"]) and np.array_equal(depths, g["depths"])
    ch, _ = env.expand(st)
    assert _sha(ch) == str(g["children_sha256"])
    assert np.array_equal(np.packbits(env.is_solved(ch.reshape(-1, dim * dim)).reshape(2000, 4)), g["solved"])


@pytest.mark.parametrize("dim", [4, 5, 6, 7])
def test_swap_table_restatement(dim):
    assert np.array_equal(O.OracleNPuzzle(dim).swap, O.OracleNPuzzle.build_swap_table(dim))


@pytest.mark.parametrize("name", ["cube3", "puzzle15", "puzzle48"])
def test_golden_triples_replay(golden_dir, name):
    """21,349 / 26,011 / 127,835 (s, a, s') triples of the reference's shipped BWAS results."""
    g = np.load(golden_dir + "/paths_%s.npz" % name)
    env = O.get_oracle_env(name)
    states, moves, offs = g["states"], g["moves"], g["offsets"]
    is_last = np.zeros(len(states), bool); is_last[offs[1:] - 1] = True
    src, dst = states[~is_last], states[np.roll(~is_last, 1)]
    assert len(src) == len(moves) == {"cube3": 21349, "puzzle15": 26011, "puzzle48": 127835}[name]
    for a in np.unique(moves):
        m = moves == a
        assert np.array_equal(env.move(src[m], int(a)), dst[m])
    assert np.array_equal(env.is_solved(states), is_last)


@pytest.mark.parametrize("name", ["cube3", "puzzle15"])
def test_shipped_optimal_solutions_solve(golden_dir, name):
    g = np.load(golden_dir + "/optimal_%s.npz" % name)
    env = O.get_oracle_env(name)
    st, moves, offs = g["states"].copy(), g["moves"], g["offsets"]
    lens = np.diff(offs)
    for k in range(int(lens.max())):
        act = lens > k
        mv = moves[offs[:-1][act] + k]
        for a in np.unique(mv):
            rows = np.where(act)[0][mv == a]
            st[rows] = env.move(st[rows], int(a))
    assert env.is_solved(st).all()


def test_cube4_oracle_reproduces_reference(golden_dir):
    """4x4x4 cube (C++-only in the reference): the 24 permutations, all children and Cube4::isSolved of 1564
    parents -- scrambles, whole-cube rotations, stickers exchanged inside a face and their neighbours -- as dumped from the
    compiled reference class (tests/golden/make_golden_cube4.py)."""
    env = O.get_oracle_env("cube4")
    t = json.load(open(golden_dir + "/cube4_tables.json"))
    perm = np.array(t["perm"])
    assert perm.shape == (24, 96) and all(sorted(r) == list(range(96)) for r in perm.tolist())
    assert [int((r != np.arange(96)).sum()) for r in perm] == [32] * 12 + [16] * 12        # outer layers, inner slices
    g = np.load(golden_dir + "/cube4_cfg1.npz")
    ch, _ = env.expand(g["parents"])
    assert _sha(ch) == str(g["children_sha256"]) and np.array_equal(ch[:64], g["children_head"])
    sol = np.concatenate([env.is_solved(g["parents"])[:, None], env.is_solved(ch.reshape(-1, 96)).reshape(-1, 24)], axis=1)
    assert np.array_equal(sol, g["solved"].astype(bool)) and sol.sum() == int(g["n_solved"]) > 100
    assert sol[:, 0].sum() > 10 and not np.array_equal(g["parents"][sol[:, 0]][-1], env.goal)  # solved, yet not the identity state
    for a in range(24):
        assert np.array_equal(env.prev(env.move(g["parents"], a), a), g["parents"])


@pytest.mark.parametrize("name", ["cube3", "cube4", "puzzle15", "puzzle24", "puzzle35", "puzzle48"])
def test_c_oracle_equals_numpy_oracle(oracle_clib, golden_dir, name):
    env = O.get_oracle_env(name)
    np.random.seed(4); random.seed(4)
    st, _ = env.generate_states(3000, (0, 20))
    n, S, A = len(st), env.state_dim, env.num_moves
    ch = np.empty((n, A, S), np.uint8); sv = np.empty((n, A), np.uint8); hs = np.empty(n * A, np.uint64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    if name == "cube3":
        t = json.load(open(golden_dir + "/cube3_tables.json"))
        new, old = np.array(t["idxs_new"], np.int32), np.array(t["idxs_old"], np.int32)
        oracle_clib.oracle_cube3_expand(p(st), ctypes.c_int64(n), p(new), p(old), p(ch), p(sv))
    elif name == "cube4":
        perm = np.ascontiguousarray(env.perm, dtype=np.int32)
        oracle_clib.oracle_cube4_expand(p(st), ctypes.c_int64(n), p(perm), p(ch), p(sv))
    else:
        sw = np.ascontiguousarray(env.swap, dtype=np.int32)
        oracle_clib.oracle_puzzle_expand(p(st), ctypes.c_int64(n), env.dim, p(sw), p(ch), p(sv))
    och, _ = env.expand(st)
    assert np.array_equal(ch, och)
    assert np.array_equal(sv.astype(bool).reshape(-1), env.is_solved(och.reshape(-1, S)))
    keys = O.HASH_KEYS.astype(np.uint32)
    oracle_clib.oracle_hash64(p(och.reshape(-1, S)), ctypes.c_int64(n * A), S, p(keys), ctypes.c_uint64(int(O.HASH_SEED)), p(hs))
    assert np.array_equal(hs, O.state_hash64(och.reshape(-1, S)))


def test_hash_is_injective_on_fixture_states(golden_dir):
    g = np.load(golden_dir + "/cube3_cfg1.npz")
    ch, _ = O.OracleCube3().expand(g["parents"])
    flat = ch.reshape(-1, 54)
    h = O.state_hash64(flat)
    assert (h != 0).all()
    assert len(np.unique(h)) == len(np.unique(flat, axis=0))


def test_lightsout_oracle_reproduces_reference(golden_dir, oracle_clib):
    """Lights Out 7x7: tables, seeded scrambles, all 49 children, solved flags, the reference's shipped
    (s, a, s') triples, and the C restatement."""
    env = O.OracleLightsOut(7)
    t = json.load(open(golden_dir + "/lightsout_tables.json"))
    assert np.array_equal(env.move_matrix, np.array(t["move_matrix"]))
    g = np.load(golden_dir + "/lightsout7_cfg1.npz")
    np.random.seed(4); random.seed(4)
    st, depths = env.generate_states(2000, (0, 50))
    assert np.array_equal(st, g["parents"]) and np.array_equal(depths, g["depths"])
    ch, _ = env.expand(st)
    assert _sha(ch) == str(g["children_sha256"]) and np.array_equal(ch[:64], g["children_head"])
    solved = env.is_solved(ch.reshape(-1, 49)).reshape(2000, 49)
    assert np.array_equal(np.packbits(solved), g["solved"]) and solved.sum() == int(g["n_solved"])
    pg = np.load(golden_dir + "/paths_lightsout7.npz")
    states, moves, offs = pg["states"], pg["moves"], pg["offsets"]
    is_last = np.zeros(len(states), bool); is_last[offs[1:] - 1] = True
    src, dst = states[~is_last], states[np.roll(~is_last, 1)]
    assert len(src) == len(moves) == 12130
    assert np.array_equal(env.move_many(src, moves), dst)
    assert np.array_equal(env.is_solved(states), is_last)
    n = 500
    c2 = np.empty((n, 49, 49), np.uint8); s2 = np.empty((n, 49), np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    mm = np.ascontiguousarray(env.move_matrix, dtype=np.int32)
    oracle_clib.oracle_lightsout_expand(p(st[:n]), ctypes.c_int64(n), 7, p(mm), p(c2), p(s2))
    assert np.array_equal(c2, ch[:n]) and np.array_equal(s2.astype(bool), solved[:n])
