"""GPU parity: the CUDA environment step (through the C ABI) vs the oracle, bit-exact."""
import hashlib
import random

import numpy as np
import pytest

from oracle import oracle_env as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ENVS = ["cube3", "puzzle15", "puzzle24", "puzzle35", "puzzle48", "lightsout7", "cube4"]


def _states(name, n, seed, back):
    env = O.get_oracle_env(name)
    np.random.seed(seed); random.seed(seed)
    st, _ = env.generate_states(n, back)
    return env, st


@pytest.mark.parametrize("name", ENVS)
@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 1000, 4097])
def test_expand_matches_oracle(name, n):
    from deepcubea_b200 import ops, _lib
    env, st = _states(name, max(n, 1), 5, (0, 12))
    st = st[:n]
    ch, sv, hs = ops.expand(_lib.ENV_IDS[name], torch.from_numpy(st).cuda())
    torch.cuda.synchronize()
    och, _ = env.expand(st) if n else (np.zeros((0, env.num_moves, env.state_dim), np.uint8), None)
    assert np.array_equal(ch.cpu().numpy(), och)
    flat = och.reshape(-1, env.state_dim)
    assert np.array_equal(sv.cpu().numpy().reshape(-1).astype(bool), env.is_solved(flat) if n else np.zeros(0, bool))
    assert np.array_equal(hs.cpu().numpy().reshape(-1).view(np.uint64), O.state_hash64(flat) if n else np.zeros(0, np.uint64))


def test_cube3_config1_golden(golden_dir):
    """BASELINE config 1: 10k seeded scrambles x 12 moves, digest of the reference's own output."""
    from deepcubea_b200 import ops
    g = np.load(golden_dir + "/cube3_cfg1.npz")
    par = g["parents"]
    ch, sv, _ = ops.expand(0, torch.from_numpy(par).cuda())
    ch = ch.cpu().numpy()
    assert hashlib.sha256(ch.tobytes()).hexdigest() == str(g["children_sha256"])
    assert np.array_equal(ch[:256], g["children_head"])
    assert np.array_equal(np.packbits(sv.cpu().numpy().astype(bool)), g["solved"])
    nn_in = ops.nnet_input(0, torch.from_numpy(par).cuda()).cpu().numpy()
    assert hashlib.sha256(nn_in.tobytes()).hexdigest() == str(g["nnet_in_sha256"])


@pytest.mark.parametrize("name", ["puzzle15", "puzzle48"])
def test_puzzle_config1_golden(golden_dir, name):
    from deepcubea_b200 import ops, _lib
    g = np.load(golden_dir + "/%s_cfg1.npz" % name)
    ch, sv, _ = ops.expand(_lib.ENV_IDS[name], torch.from_numpy(g["parents"]).cuda())
    ch = ch.cpu().numpy()
    assert hashlib.sha256(ch.tobytes()).hexdigest() == str(g["children_sha256"])
    assert np.array_equal(np.packbits(sv.cpu().numpy().astype(bool)), g["solved"])


def test_lightsout_config1_golden(golden_dir):
    from deepcubea_b200 import ops
    g = np.load(golden_dir + "/lightsout7_cfg1.npz")
    ch, sv, _ = ops.expand(5, torch.from_numpy(g["parents"]).cuda())
    ch = ch.cpu().numpy()
    assert hashlib.sha256(ch.tobytes()).hexdigest() == str(g["children_sha256"])
    assert np.array_equal(ch[:64], g["children_head"])
    assert np.array_equal(np.packbits(sv.cpu().numpy().astype(bool)), g["solved"])


def test_cube4_config1_golden(golden_dir):
    """4x4x4 cube: children and Cube4::isSolved flags dumped from the compiled reference class (scrambles, whole-cube
    rotations, stickers exchanged inside a face and their neighbours); the indexed kernel is covered by tests/test_gpu_bwas.py."""
    from deepcubea_b200 import ops
    g = np.load(golden_dir + "/cube4_cfg1.npz")
    par = torch.from_numpy(g["parents"]).cuda()
    ch, sv, hs = ops.expand(6, par)
    chn = ch.cpu().numpy()
    assert hashlib.sha256(chn.tobytes()).hexdigest() == str(g["children_sha256"])
    assert np.array_equal(chn[:64], g["children_head"])
    root = ops.is_solved(6, par).cpu().numpy()
    assert np.array_equal(np.concatenate([root[:, None], sv.cpu().numpy()], axis=1), g["solved"])
    assert np.array_equal(hs.cpu().numpy().reshape(-1).view(np.uint64), O.state_hash64(chn.reshape(-1, 96)))
    assert np.array_equal(ops.nnet_input(6, par).cpu().numpy(), g["parents"] // 16)


@pytest.mark.parametrize("name", ["cube3", "puzzle15", "puzzle48", "lightsout7"])
def test_golden_triples(golden_dir, name):
    """Every (s, a, s') of the reference's shipped BWAS results replays bit-exactly through next_state."""
    from deepcubea_b200 import ops, _lib
    g = np.load(golden_dir + "/paths_%s.npz" % name)
    states, moves, offs = g["states"], g["moves"], g["offsets"]
    is_last = np.zeros(len(states), bool); is_last[offs[1:] - 1] = True
    src = states[~is_last]; dst = states[np.roll(~is_last, 1)]
    assert len(src) == len(moves)
    eid = _lib.ENV_IDS[name]
    for a in np.unique(moves):
        m = moves == a
        out = ops.next_state(eid, torch.from_numpy(src[m]).cuda(), int(a)).cpu().numpy()
        assert np.array_equal(out, dst[m])
    sv = ops.is_solved(eid, torch.from_numpy(states).cuda()).cpu().numpy().astype(bool)
    assert np.array_equal(sv, is_last)      # finals solved, no intermediate state solved


@pytest.mark.parametrize("name", ENVS)
def test_single_state_ops(name):
    from deepcubea_b200 import ops, _lib
    env, st = _states(name, 777, 9, (0, 6))
    eid = _lib.ENV_IDS[name]
    d = torch.from_numpy(st).cuda()
    for a in range(env.num_moves):
        assert np.array_equal(ops.next_state(eid, d, a).cpu().numpy(), env.move(st, a))
        back = ops.next_state(eid, ops.next_state(eid, d, a), env.rev_action[a]).cpu().numpy()
        if name in ("cube3", "lightsout7", "cube4"):
            assert np.array_equal(back, st)       # move then inverse = identity
    assert np.array_equal(ops.is_solved(eid, d).cpu().numpy().astype(bool), env.is_solved(st))
    assert np.array_equal(ops.hash_states(eid, d).cpu().numpy().view(np.uint64), O.state_hash64(st))
    assert np.array_equal(ops.nnet_input(eid, d).cpu().numpy(), env.nnet_input(st))


def test_host_buffer_abi():
    """The host-pointer entry point a ctypes/cgo-style binding would call."""
    import ctypes
    from deepcubea_b200 import _lib
    lib = _lib.load()
    env, st = _states("cube3", 5000, 3, (0, 10))
    st = np.ascontiguousarray(st)
    ch = np.empty((5000, 12, 54), np.uint8); sv = np.empty((5000, 12), np.uint8); hs = np.empty((5000, 12), np.uint64)
    _lib.check(lib.dcb_expand_host(0, _lib.ptr(st), 5000, _lib.ptr(ch), _lib.ptr(sv), _lib.ptr(hs), 0))
    och, _ = env.expand(st)
    assert np.array_equal(ch, och)
    assert np.array_equal(hs.reshape(-1), O.state_hash64(och.reshape(-1, 54)))


def test_expand_full_size_properties():
    """BASELINE-size property checks (no oracle loop): 4 quarter turns = identity; move o inverse = identity;
    children hashes equal hashes recomputed from the children."""
    from deepcubea_b200 import ops
    n = 1 << 20
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    st = torch.arange(54, dtype=torch.uint8, device="cuda").repeat(n, 1)
    acts = torch.randint(0, 12, (30,), generator=g, device="cuda").tolist()
    for a in acts:                       # same scramble for all, then diversify by index
        st = ops.next_state(0, st, a)
    idx = torch.arange(n, device="cuda")
    for a in range(12):
        sel = (idx % 12) == a
        st[sel] = ops.next_state(0, st[sel].contiguous(), a)
    ch, sv, hs = ops.expand(0, st)
    assert int(sv.sum()) == 0
    for a in (0, 5, 11):
        c = ch[:, a].contiguous()
        assert torch.equal(ops.hash_states(0, c), hs[:, a].contiguous())
        c4 = c
        for _ in range(3):
            c4 = ops.next_state(0, c4, a)
        assert torch.equal(c4, st)       # a^4 = identity
