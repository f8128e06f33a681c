"""CPU tier: host-side logic of the product (no GPU compute): table construction, device math compiled for
the host vs the oracle, C-ABI surface, plugin-API plumbing, pickles, CLI parsing."""
import ctypes
import json
import os
import pickle
import random
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle_env as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_geometric_tables_equal_reference(golden_dir):
    from deepcubea_b200.environments import cube3_geometry as G
    t = json.load(open(golden_dir + "/cube3_tables.json"))
    assert np.array_equal(G.move_permutations(), np.array(t["perm"]))
    assert G.MOVES == t["moves"] and G.MOVES_REV == t["moves_rev"]
    assert G.inverse_actions() == [1, 0, 3, 2, 5, 4, 7, 6, 9, 8, 11, 10]


def test_generated_prmt_header_is_current():
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "deepcubea_b200", "csrc", "gen_cube3_moves.py"), "--check"],
                         stdout=subprocess.DEVNULL)
    assert rc == 0


@pytest.mark.parametrize("envid,name", list(enumerate(["cube3", "puzzle15", "puzzle24", "puzzle35", "puzzle48"])))
def test_device_math_on_host_matches_oracle(hostcheck_lib, envid, name):
    """The kernels' register-level math (PRMT networks, SIMD blank swap, record packing, hash, is_solved),
    compiled with g++ from the same headers, against the oracle."""
    env = O.get_oracle_env(name)
    np.random.seed(7); random.seed(7)
    st, _ = env.generate_states(4000, (0, 9))
    n, S, A = len(st), env.state_dim, env.num_moves
    buf = np.zeros(n * S + 8, np.uint8); buf[:n * S] = st.reshape(-1)
    ch = np.zeros((n, A, S), np.uint8); sv = np.zeros((n, A), np.uint8); hs = np.zeros((n, A), np.uint64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert hostcheck_lib.hc_expand(envid, p(buf), ctypes.c_int64(n), p(ch), p(sv), p(hs)) == 0
    och, _ = env.expand(st)
    assert np.array_equal(ch, och)
    assert np.array_equal(sv.astype(bool).reshape(-1), env.is_solved(och.reshape(-1, S)))
    assert sv.sum() > 0
    assert np.array_equal(hs.reshape(-1), O.state_hash64(och.reshape(-1, S)))
    for a in range(A):
        out = np.zeros((n, S), np.uint8)
        assert hostcheck_lib.hc_next_state(envid, p(buf), ctypes.c_int64(n), a, p(out)) == 0
        assert np.array_equal(out, env.move(st, a))


def test_cube4_geometry_equals_reference_tables(golden_dir):
    """24 x 96 permutation built from the 3-D embedding == the children of the identity state dumped from the compiled
    reference Cube4 class (tests/golden/make_golden_cube4.py)."""
    from deepcubea_b200.environments import cube4_geometry as G
    t = json.load(open(golden_dir + "/cube4_tables.json"))
    assert np.array_equal(G.move_permutations(), np.array(t["perm"]))
    assert G.inverse_actions() == [a ^ 1 for a in range(24)] and len(G.MOVES) == 24


def test_cube4_device_math_on_host_matches_oracle_and_reference(hostcheck_lib, golden_dir):
    """cube4 PRMT networks, hash and the colour-uniformity solved test (compiled for the host from the kernels' headers) on the
    golden parents: children sha256 and Cube4::isSolved flags of the reference itself, and the oracle on fresh scrambles."""
    import hashlib
    env = O.get_oracle_env("cube4")
    g = np.load(golden_dir + "/cube4_cfg1.npz")
    np.random.seed(5); random.seed(5)
    st2, _ = env.generate_states(1500, (0, 6))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for st, golden in ((g["parents"], True), (st2, False)):
        n = len(st)
        buf = np.zeros(n * 96 + 8, np.uint8); buf[:n * 96] = st.reshape(-1)
        ch = np.zeros((n, 24, 96), np.uint8); sv = np.zeros((n, 24), np.uint8); hs = np.zeros((n, 24), np.uint64)
        assert hostcheck_lib.hc_expand(6, p(buf), ctypes.c_int64(n), p(ch), p(sv), p(hs)) == 0
        root = np.zeros(n, np.uint8)
        assert hostcheck_lib.hc_is_goal(6, p(buf), ctypes.c_int64(n), p(root)) == 0
        if golden:
            assert hashlib.sha256(ch.tobytes()).hexdigest() == str(g["children_sha256"])
            assert np.array_equal(np.concatenate([root[:, None], sv], 1), g["solved"]) and g["solved"].sum() > 100
        och, _ = env.expand(st)
        assert np.array_equal(ch, och)
        assert np.array_equal(sv.astype(bool).reshape(-1), env.is_solved(och.reshape(-1, 96)))
        assert np.array_equal(root.astype(bool), env.is_solved(st))
        assert np.array_equal(hs.reshape(-1), O.state_hash64(och.reshape(-1, 96)))
    for a in range(24):
        out = np.zeros((len(st2), 96), np.uint8)
        buf = np.zeros(len(st2) * 96 + 8, np.uint8); buf[:len(st2) * 96] = st2.reshape(-1)
        assert hostcheck_lib.hc_next_state(6, p(buf), ctypes.c_int64(len(st2)), a, p(out)) == 0
        assert np.array_equal(out, env.move(st2, a))


def test_lightsout_device_math_on_host_matches_oracle(hostcheck_lib):
    env = O.OracleLightsOut(7)
    np.random.seed(9); random.seed(9)
    st, _ = env.generate_states(600, (0, 6))
    n = len(st)
    buf = np.zeros(n * 49 + 8, np.uint8); buf[:n * 49] = st.reshape(-1)
    ch = np.zeros((n, 49, 49), np.uint8); sv = np.zeros((n, 49), np.uint8); hs = np.zeros((n, 49), np.uint64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert hostcheck_lib.hc_lightsout_expand(p(buf), ctypes.c_int64(n), p(ch), p(sv), p(hs)) == 0
    och, _ = env.expand(st)
    assert np.array_equal(ch, och)
    assert np.array_equal(sv.astype(bool).reshape(-1), env.is_solved(och.reshape(-1, 49))) and sv.sum() > 0
    assert np.array_equal(hs.reshape(-1), O.state_hash64(och.reshape(-1, 49)))


def test_c_abi_exports_every_declared_symbol():
    """The library loads and exports exactly what include/dcb.h declares (no compute without a GPU)."""
    from deepcubea_b200 import _lib
    header = open(os.path.join(ROOT, "include", "dcb.h")).read()
    declared = sorted(set(re.findall(r"\b(dcb_[a-z0-9_]+)\s*\(", header)) - {"dcb_open_state"})
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == _lib.exported_symbols()
    assert lib.dcb_abi_version() == 1
    assert [lib.dcb_env_state_bytes(e) for e in range(6)] == [54, 16, 25, 36, 49, 49]
    assert [lib.dcb_env_num_moves(e) for e in range(6)] == [12, 4, 4, 4, 4, 49]
    assert [lib.dcb_env_slot_align(e) for e in range(6)] == [2, 1, 4, 1, 4, 16]
    assert lib.dcb_env_state_bytes(9) == -1 and lib.dcb_error_string(-1) == b"unknown environment id"
    assert lib.dcb_expand(9, None, 0, None, None, None, None) == -1            # argument validation happens before any launch
    assert lib.dcb_expand(0, None, 5, None, None, None, None) == -2
    assert lib.dcb_closed_bytes(1 << 20) == 16 << 20 and lib.dcb_closed_bytes(1000) < 0


def test_abi_tables_match_reference(golden_dir):
    from deepcubea_b200 import _lib
    lib = _lib.load()
    t = json.load(open(golden_dir + "/cube3_tables.json"))
    buf = np.zeros(12 * 54, np.int32)
    assert lib.dcb_env_move_table(0, _lib.ptr(buf), buf.size) == 0
    assert np.array_equal(buf.reshape(12, 54), np.array(t["perm"]))
    pz = json.load(open(golden_dir + "/puzzle_tables.json"))
    for e, dim in ((1, 4), (2, 5), (3, 6), (4, 7)):
        buf = np.zeros(dim * dim * 4, np.int32)
        assert lib.dcb_env_move_table(e, _lib.ptr(buf), buf.size) == 0
        assert np.array_equal(buf.reshape(-1, 4), np.array(pz[str(dim)]["swap_zero_idxs"]))
        goal = np.zeros(dim * dim, np.uint8)
        assert lib.dcb_env_goal_state(e, _lib.ptr(goal)) == 0
        assert np.array_equal(goal, np.array(pz[str(dim)]["goal"]))
    lo = json.load(open(golden_dir + "/lightsout_tables.json"))
    buf = np.zeros(49 * 5, np.int32)
    assert lib.dcb_env_move_table(5, _lib.ptr(buf), buf.size) == 0
    assert np.array_equal(buf.reshape(49, 5), np.array(lo["move_matrix"]))
    goal = np.ones(49, np.uint8)
    assert lib.dcb_env_goal_state(5, _lib.ptr(goal)) == 0 and goal.sum() == 0
    c4 = json.load(open(golden_dir + "/cube4_tables.json"))
    buf = np.zeros(24 * 96, np.int32)
    assert lib.dcb_env_move_table(6, _lib.ptr(buf), buf.size) == 0 and lib.dcb_env_move_table(6, _lib.ptr(buf), 100) < 0
    assert np.array_equal(buf.reshape(24, 96), np.array(c4["perm"]))
    goal = np.zeros(96, np.uint8)
    assert lib.dcb_env_goal_state(6, _lib.ptr(goal)) == 0 and np.array_equal(goal, np.arange(96))
    assert (lib.dcb_env_state_bytes(6), lib.dcb_env_num_moves(6), lib.dcb_env_slot_align(6)) == (96, 24, 1)
    assert lib.dcb_env_state_bytes(7) < 0


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from deepcubea_b200._lib import DcbError
    from deepcubea_b200.utils.env_utils import get_environment
    env = get_environment("cube3")
    with pytest.raises(DcbError):
        env.is_solved(env.generate_goal_states(2))
    with pytest.raises(DcbError):
        env.expand(env.generate_goal_states(2))
    from deepcubea_b200.search.bwas_gpu import BWASGpu
    with pytest.raises(DcbError):
        BWASGpu("cube3", lambda x: x, 0.8, 10)


def test_product_never_imports_oracle():
    """Parity claims are void if the product routes through the checker."""
    for base, _, files in os.walk(os.path.join(ROOT, "deepcubea_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("see oracle/", "").replace("oracle/oracle_", "ORACLE_DOC_"), f


def test_registry_and_plugin_api_surface():
    from deepcubea_b200.environments.environment_abstract import Environment
    from deepcubea_b200.utils.env_utils import get_environment
    for name, S, A in (("cube3", 54, 12), ("puzzle15", 16, 4), ("PUZZLE24", 25, 4), ("puzzle35", 36, 4), ("puzzle48", 49, 4), ("lightsout7", 49, 49),
                       ("cube4", 96, 24)):
        env = get_environment(name)
        assert isinstance(env, Environment) and env.get_num_moves() == A and env.state_dim == S
        for m in ("next_state", "prev_state", "generate_goal_states", "is_solved", "state_to_nnet_input", "get_nnet_model",
                  "generate_states", "expand"):
            assert callable(getattr(env, m))
        goals = env.generate_goal_states(3)
        assert len({hash(g) for g in goals}) == 1 and goals[0] == goals[1]
        assert env.generate_goal_states(2, np_format=True).shape == (2, S)
    for bad in ("lightsout5", "sokoban", "cube5"):
        with pytest.raises(ValueError):
            get_environment(bad)
    env = get_environment("puzzle15")
    assert np.array_equal(env.swap_zero_idxs, O.OracleNPuzzle(4).swap)
    assert np.array_equal(get_environment("lightsout7").move_matrix, O.OracleLightsOut(7).move_matrix)
    model = get_environment("cube3").get_nnet_model()
    assert sum(p.numel() for p in model.parameters()) == 14_688_001 or sum(p.numel() for p in model.parameters()) > 14_600_000


def test_state_classes_pickle_with_reference_module_paths(tmp_path):
    sys.path.insert(0, ROOT)
    import environments.cube3 as rc
    import environments.n_puzzle as rn
    from deepcubea_b200.environments.cube3 import Cube3State
    from deepcubea_b200.environments.n_puzzle import NPuzzleState
    assert rc.Cube3State is Cube3State and rn.NPuzzleState is NPuzzleState
    s = Cube3State(np.arange(54, dtype=np.uint8))
    blob = pickle.dumps({"states": [s], "paths": [[s, s]]}, protocol=-1)
    assert b"environments.cube3" in blob and b"deepcubea_b200" not in blob
    back = pickle.loads(blob)["states"][0]
    assert back == s and hash(back) == hash(s)
    # shipped pickles hold int64 payloads and may leave the `hash` slot unset
    t = NPuzzleState.__new__(NPuzzleState); t.tiles = np.arange(16, dtype=np.int64)
    assert isinstance(hash(t), int)
    from deepcubea_b200.utils.env_utils import get_environment
    assert get_environment("puzzle15").pack([t]).dtype == np.uint8


def test_misc_utils_and_logger(tmp_path, capsys):
    from deepcubea_b200.utils import data_utils, misc_utils
    data = [[1, 2], [], [3], [4, 5, 6]]
    flat, idx = misc_utils.flatten(data)
    assert flat == [1, 2, 3, 4, 5, 6] and misc_utils.unflatten(flat, idx) == data
    assert misc_utils.split_evenly(10, 4) == [3, 3, 2, 2]
    lg = data_utils.Logger(str(tmp_path / "o.txt"), "w")
    lg.write("hello\n"); lg.flush()
    assert open(tmp_path / "o.txt").read() == "hello\n"


def test_astar_cli_flags_match_reference():
    from deepcubea_b200.search_methods import astar
    with pytest.raises(SystemExit):
        astar.main(["--help"])
    with pytest.raises(SystemExit):          # the three required flags of astar.py:346-353
        astar.main(["--env", "cube3"])
    src = open(os.path.join(ROOT, "deepcubea_b200", "search_methods", "astar.py")).read()
    for flag in ("--states", "--model_dir", "--env", "--batch_size", "--weight", "--language", "--results_dir", "--start_idx",
                 "--nnet_batch_size", "--verbose", "--debug"):
        assert flag in src


def test_get_path_and_node_api():
    from deepcubea_b200.search_methods.astar import Node, get_path
    a = Node("s0", 0.0, False, None, None); b = Node("s1", 1.0, False, 3, a); c = Node("s2", 2.0, True, 7, b)
    path, moves, cost = get_path(c)
    assert path == ["s0", "s1", "s2"] and moves == [3, 7] and cost == 2.0


def test_folded_network_equals_reference_module_fp32():
    """BN folding (fp64 fold, fp32 store) vs the unfolded module, CPU fp32: same function to ~1e-5."""
    import torch
    from deepcubea_b200.nnet.folded import FoldedResnet
    from deepcubea_b200.utils.pytorch_models import ResnetModel
    torch.manual_seed(0)
    m = ResnetModel(16, 16, 64, 32, 2, 1, True)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.normal_(0, 0.3); mod.running_var.uniform_(0.5, 2.0); mod.weight.data.uniform_(0.5, 1.5); mod.bias.data.normal_()
    m.eval()
    x = torch.randint(0, 16, (257, 16), dtype=torch.uint8)
    with torch.no_grad():
        ref = m(x)[:, 0]
    got = FoldedResnet(m, "fp32")(x)
    assert torch.allclose(ref, got, atol=2e-5, rtol=1e-5)


def test_closed_four_phase_rule_equals_the_sequential_loop():
    """The CLOSED kernels (closed_table.cu) settle a whole batch with four data-parallel phases -- probe (record the pre-batch value),
    min (fold (g, id) of the candidates that beat it), resolve (winner / earlier-and-cheaper winner / left to the fix-up), fix-up
    (running minimum among the few that came before their state's winner).  Pure-Python model of those phases against the
    reference's child-order loop (parallel_weighted_astar.cpp:246-261; remove_in_closed, astar.py:78-90) on random batches with
    many repeated states and depths."""
    rng = np.random.RandomState(7)
    table = {}                                            # state -> (g, id): what the slots hold
    seq = {}                                              # the reference's dict: state -> best g
    next_id = 0
    for rnd in range(40):
        m = int(rng.randint(1, 400))
        states = rng.randint(0, 60, m)                    # few distinct states: many in-batch duplicates
        g = rng.randint(1, 9, m)
        ids = np.arange(next_id, next_id + m)
        next_id += m
        # reference: sequential in child order
        keep_ref = np.zeros(m, bool)
        for i in range(m):
            if states[i] not in seq or seq[states[i]] > g[i]:
                seq[states[i]] = g[i]; keep_ref[i] = True
        # phase 1: probe -- every candidate records the slot's value BEFORE the batch
        prev = [table.get(s) for s in states]
        # phase 2: min -- candidates that beat the recorded value fold (g, id) into the slot
        for i in range(m):
            if prev[i] is None or g[i] < prev[i][0]:
                cur = table.get(states[i])
                if cur is None or (g[i], ids[i]) < cur:
                    table[states[i]] = (g[i], ids[i])
        # phase 3: resolve
        keep = np.zeros(m, bool)
        amb = []
        for i in range(m):
            if prev[i] is not None and g[i] >= prev[i][0]:
                continue                                  # not cheaper than what CLOSED held: dropped
            win = table[states[i]]
            if win == (g[i], ids[i]):
                keep[i] = True
            elif win[1] > ids[i]:
                amb.append(i)                             # came BEFORE the winner with a larger g
        # phase 4: fix-up -- kept iff no earlier candidate of the same state on the list is at least as cheap
        for i in amb:
            if not any(states[j] == states[i] and ids[j] < ids[i] and g[j] <= g[i] for j in amb):
                keep[i] = True
        assert np.array_equal(keep, keep_ref), "round %d" % rnd
        assert {s: v[0] for s, v in table.items()} == seq


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """The Python mirrors of the ABI structs (deepcubea_b200/_lib.py) against include/dcb.h compiled with gcc: sizes and the offsets
    of every field a binding touches."""
    import ctypes
    import subprocess
    from deepcubea_b200 import _lib
    fields = {"dcb_search_inst": ["open_size", "n_popped", "n_expand", "goal_id", "goal_key", "done", "n_goals", "next_slot", "iterations",
                                  "nodes_generated", "nodes_expanded", "thr_key", "n_take", "resting", "overflow", "thr_lo"],
              "dcb_step_plan": ["n_tiles", "n_parents", "n_kept", "closed_entries", "n_running", "error", "budget", "total_kept", "total_expanded"],
              "dcb_search_ctx": ["env", "semantics", "slots_per_inst", "open_per_inst", "closed_capacity", "d_arena", "d_closed", "d_open_key",
                                 "d_open_id", "d_open_key_lo", "d_inst", "d_plan", "d_weights", "d_kept_ids", "d_closed_scratch"]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dcb.h"', 'int main(void) {']
    for st, fs in fields.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (st, st))
        for f in fs:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (st, f, st, f))
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    mirrors = {"dcb_search_inst": _lib.SearchInst, "dcb_step_plan": _lib.StepPlan, "dcb_search_ctx": _lib.SearchCtx}
    for st, fs in fields.items():
        assert int(got[st]) == ctypes.sizeof(mirrors[st]), st
        for f in fs:
            assert int(got["%s.%s" % (st, f)]) == getattr(mirrors[st], f).offset, "%s.%s" % (st, f)
    assert ctypes.sizeof(_lib.SearchInst) == 128 and ctypes.sizeof(_lib.StepPlan) == 64


def test_bench_arms_share_one_state_list_and_interval_union():
    """bench.py: both arms draw the SAME seeded start states (the reference arm may run on a GPU-less host: the checker's numpy port
    of the reference generator gives the list the GPU-backed environment gives, tests/test_gpu_cli.py), scrambles of depth >= 20 only,
    prefix-stable in the number asked for; and the union-of-intervals helper behind the GEMM busy time."""
    sys.path.insert(0, ROOT)
    import bench
    a, desc_a = bench.workload_states("cube3", 8)
    b, desc_b = bench.workload_states("cube3", 24)
    assert a.shape == (8, 54) and b.shape == (24, 54) and np.array_equal(a, b[:8])          # rank r of N takes b[r::N]: same list for every N
    assert "seed %d" % bench.SEED in desc_a and "depth >= %d" % bench.MIN_SCRAMBLE in desc_a
    env = O.OracleCube3()
    np.random.seed(bench.SEED); random.seed(bench.SEED)
    st, depths = env.generate_states(bench.POOL, (0, 26))
    keep = np.nonzero(np.asarray(depths) >= bench.MIN_SCRAMBLE)[0]
    assert np.array_equal(a, st[keep[:8]])
    assert bench.interval_union([]) == 0.0
    assert bench.interval_union([(0.0, 1.0), (2.0, 3.0)]) == 2.0
    assert bench.interval_union([(0.0, 2.0), (1.0, 3.0), (2.5, 2.75), (5.0, 6.0)]) == 4.0
