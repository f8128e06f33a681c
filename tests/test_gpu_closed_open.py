"""GPU: CLOSED (hash table) and OPEN (radix-select queue) through the C ABI against plain-Python models of the
reference's containers (std::unordered_set + depth rule, std::priority_queue + pop loop)."""
import ctypes
import heapq
import random

import numpy as np
import pytest

from oracle import oracle_env as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _u32(t):
    return t.cpu().numpy().view(np.uint32)


def _scratch(lib, m):
    return torch.empty(int(lib.dcb_closed_scratch_bytes(m)), dtype=torch.uint8, device="cuda")


def test_closed_insert_or_improve_matches_dict_model():
    from deepcubea_b200 import _lib
    lib = _lib.load(); p = _lib.ptr
    env = O.OracleCube3()
    rng = np.random.RandomState(0); random.seed(0)
    np.random.seed(0)
    pool, _ = env.generate_states(3000, (0, 7))              # many repeated states at shallow depth
    cap = 1 << 14
    table = torch.empty(cap * 2, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.dcb_closed_clear(p(table), cap, st))
    counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    arena = torch.zeros(20000 * 54 + 64, dtype=torch.uint8, device="cuda")
    model = {}                                               # state bytes -> best g
    next_id = 0
    for rnd in range(6):
        m = 1500
        idx = rng.randint(0, len(pool), m)
        states = pool[idx]
        g = rng.randint(1, 6, m).astype(np.uint32)
        arena[next_id * 54:(next_id + m) * 54] = torch.from_numpy(states.reshape(-1)).cuda()
        hs = torch.from_numpy(O.state_hash64(states).view(np.int64)).cuda()
        gd = torch.from_numpy(g.view(np.int32)).cuda()
        slot = _scratch(lib, m); keep = torch.empty(m, dtype=torch.uint8, device="cuda")
        _lib.check(lib.dcb_closed_insert(0, p(table), cap, p(arena), p(hs), p(gd), None, next_id, m, p(slot), p(keep), p(counter), st))
        keep = keep.cpu().numpy().astype(bool)
        # model: the reference's child-order loop (parallel_weighted_astar.cpp:246-261; remove_in_closed, astar.py:78-90) -- kept
        # iff unseen or strictly cheaper than what the table holds when the loop reaches the candidate
        exp = np.zeros(m, bool)
        for i in range(m):
            k = states[i].tobytes()
            if k not in model or model[k] > g[i]:
                model[k] = g[i]; exp[i] = True
        assert np.array_equal(keep, exp), "round %d" % rnd
        next_id += m
    assert int(counter.cpu()[0]) == len(model)               # distinct states == occupied slots
    # rehash into a table 4x larger: same contents -> re-inserting everything with the same g keeps nothing
    cap2 = cap * 4
    table2 = torch.empty(cap2 * 2, dtype=torch.int64, device="cuda")
    _lib.check(lib.dcb_closed_clear(p(table2), cap2, st))
    _lib.check(lib.dcb_closed_rehash(p(table), cap, p(table2), cap2, st))
    keys = list(model.keys())
    states = np.stack([np.frombuffer(k, dtype=np.uint8) for k in keys])
    m = len(keys)
    arena2 = torch.cat([arena[:next_id * 54], torch.from_numpy(states.reshape(-1)).cuda(), torch.zeros(64, dtype=torch.uint8, device="cuda")])
    hs = torch.from_numpy(O.state_hash64(states).view(np.int64)).cuda()
    gd = torch.from_numpy(np.array([model[k] for k in keys], np.uint32).view(np.int32)).cuda()
    slot = _scratch(lib, m); keep = torch.empty(m, dtype=torch.uint8, device="cuda")
    _lib.check(lib.dcb_closed_insert(0, p(table2), cap2, p(arena2), p(hs), p(gd), None, next_id, m, p(slot), p(keep), None, st))
    assert int(keep.sum()) == 0
    gd2 = (gd - 1).clamp(min=0)                               # strictly smaller g re-opens (parallel_weighted_astar.cpp:255-261)
    _lib.check(lib.dcb_closed_insert(0, p(table2), cap2, p(arena2), p(hs), p(gd2), None, next_id, m, p(slot), p(keep), None, st))
    assert np.array_equal(keep.cpu().numpy().astype(bool), (gd2 < gd).cpu().numpy())


def test_closed_hash_collision_is_kept_not_dropped():
    """Two DIFFERENT states forced onto the same 64-bit key: the loser is verified against the arena and kept."""
    from deepcubea_b200 import _lib
    lib = _lib.load(); p = _lib.ptr
    env = O.OracleCube3()
    a = env.goal.copy(); b = env.move(a[None], 3)[0]
    arena = torch.from_numpy(np.concatenate([a, b, np.zeros(64, np.uint8)])).cuda()
    cap = 1 << 8
    table = torch.empty(cap * 2, dtype=torch.int64, device="cuda"); st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.dcb_closed_clear(p(table), cap, st))
    hs = torch.tensor([12345, 12345], dtype=torch.int64, device="cuda")           # fake equal hashes
    g = torch.tensor([2, 2], dtype=torch.int32, device="cuda")
    slot = _scratch(lib, 2); keep = torch.empty(2, dtype=torch.uint8, device="cuda")
    _lib.check(lib.dcb_closed_insert(0, p(table), cap, p(arena), p(hs), p(g), None, 0, 2, p(slot), p(keep), None, st))
    assert keep.cpu().tolist() == [1, 1]
    arena2 = torch.from_numpy(np.concatenate([a, a, np.zeros(64, np.uint8)])).cuda()   # a true duplicate is dropped
    _lib.check(lib.dcb_closed_clear(p(table), cap, st))
    _lib.check(lib.dcb_closed_insert(0, p(table), cap, p(arena2), p(hs), p(g), None, 0, 2, p(slot), p(keep), None, st))
    assert keep.cpu().tolist() == [1, 0]


@pytest.mark.parametrize("batch,stop", [(1, 0), (7, 0), (500, 0), (500, 1), (33, 1)])
def test_open_push_pop_matches_heap_model(batch, stop):
    from deepcubea_b200 import _lib
    lib = _lib.load(); p = _lib.ptr
    rng = np.random.RandomState(batch + stop)
    cap = 1 << 16
    n_nodes = 40000
    key = torch.empty(cap, dtype=torch.int32, device="cuda"); ids = torch.empty(cap, dtype=torch.int32, device="cuda")
    state = torch.zeros(_lib.INST_WORDS, dtype=torch.int32, device="cuda")
    scratch = torch.empty(int(lib.dcb_open_scratch_bytes(cap, batch)) + 16, dtype=torch.uint8, device="cuda")
    popped = torch.empty(batch, dtype=torch.int32, device="cuda")
    solved = (rng.rand(n_nodes) < 0.002).astype(np.uint8)
    solved_d = torch.from_numpy(solved).cuda()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.dcb_open_clear(p(state), st))
    heap = []
    goal = None; done = False
    next_id = 0
    for rnd in range(25):
        m = int(rng.randint(1, 1200))
        cost = (rng.randint(0, 60, m) / 4.0).astype(np.float32)                 # many exact ties
        keep = (rng.rand(m) < 0.8).astype(np.uint8)
        cd = torch.from_numpy(cost).cuda(); kd = torch.from_numpy(keep).cuda()
        _lib.check(lib.dcb_open_push(p(state), p(key), p(ids), cap, p(cd), None, next_id, p(kd), m, st))
        for i in range(m):
            if keep[i]:
                heapq.heappush(heap, (cost[i], next_id + i))
        next_id += m
        _lib.check(lib.dcb_open_pop(p(state), p(key), p(ids), cap, batch, stop, p(solved_d), p(popped), p(scratch), st))
        s = _lib.OpenState.from_buffer_copy(state.cpu().numpy().tobytes())
        # model: pop in (cost, id) order; with `stop`, break after the first solved node (parallel_weighted_astar.cpp:177-208)
        exp = []
        goal_prev = goal is not None
        for _ in range(min(batch, len(heap))):
            c, nid = heapq.heappop(heap)
            exp.append((c, nid))
            if stop and solved[nid]:
                if batch == 1:
                    goal = (c, nid); done = True
                elif goal is None or goal[0] > c:
                    goal = (c, nid)
                break
        if stop and goal_prev and exp and exp[0][0] >= goal[0]:
            done = True
        got = _u32(popped[:s.n_popped]).tolist()
        assert got == [nid for _, nid in exp], "round %d" % rnd
        assert s.size == len(heap)
        if exp:
            assert np.float32(np.array([s.min_key], np.uint32).view(np.float32)[0]) == exp[0][0]
        if stop:
            assert (s.goal_id == 0xFFFFFFFF) == (goal is None)
            if goal is not None:
                assert s.goal_id == goal[1]
            assert bool(s.done) == done
            if done:
                break
    # drain: everything comes out in (cost, id) order
    rest = []
    while True:
        _lib.check(lib.dcb_open_pop(p(state), p(key), p(ids), cap, batch, 0, None, p(popped), p(scratch), st))
        s = _lib.OpenState.from_buffer_copy(state.cpu().numpy().tobytes())
        if s.n_popped == 0:
            break
        rest += _u32(popped[:s.n_popped]).tolist()
        if len(rest) > 200000:
            break
    assert rest == [nid for _, nid in sorted(heap)]


def test_cost_and_path_kernels():
    from deepcubea_b200 import _lib
    lib = _lib.load(); p = _lib.ptr
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.RandomState(1)
    n = 5000
    h = (rng.randn(n) * 3).astype(np.float32)
    g = rng.randint(0, 40, n).astype(np.uint32); sv = (rng.rand(n) < 0.1).astype(np.uint8)
    idx = rng.permutation(n).astype(np.uint32)
    cost = torch.empty(n, dtype=torch.float32, device="cuda")
    hd, idd, gd, svd = torch.from_numpy(h).cuda(), torch.from_numpy(idx.view(np.int32)).cuda(), torch.from_numpy(g.view(np.int32)).cuda(), torch.from_numpy(sv).cuda()
    _lib.check(lib.dcb_compute_cost(p(hd), p(idd), p(gd), p(svd), 0.6, n, p(cost), st))
    w = np.float32(0.6)
    exp = (np.maximum(h, np.float32(0)) * (1 - sv[idx]).astype(np.float32)).astype(np.float32) + (w * g[idx].astype(np.float32)).astype(np.float32)
    assert np.array_equal(cost.cpu().numpy(), exp.astype(np.float32))         # bit-exact float32, no FMA contraction
    # path: chain of slots 0 <- 3 <- 9 <- 20 with moves 5, 7, 2
    A = 12
    slot_parent = np.zeros(64, np.uint32)
    ids = [0, 3 * A + 5, 9 * A + 7, 20 * A + 2]
    for child, parent in zip(ids[1:], ids[:-1]):
        slot_parent[child // A] = parent
    moves = torch.empty(16, dtype=torch.uint8, device="cuda"); ln = torch.zeros(1, dtype=torch.int32, device="cuda")
    spd = torch.from_numpy(slot_parent.view(np.int32)).cuda()
    _lib.check(lib.dcb_reconstruct_path(0, p(spd), ids[-1], 16, p(moves), p(ln), st))
    assert int(ln.cpu()[0]) == 3 and moves[:3].cpu().tolist() == [5, 7, 2]
    _lib.check(lib.dcb_reconstruct_path(0, p(spd), ids[-1], 2, p(moves), p(ln), st))
    assert int(ln.cpu()[0]) == -1
