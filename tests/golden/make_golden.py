#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz, *.json

It imports the reference from /root/reference with a numpy-2 shim (np.float/np.int were removed;
environment_abstract.py:20, n_puzzle.py:38 use them) and dumps:

  cube3_tables.json     perm[12][54] (child[j] = parent[perm[a][j]]) + the 24-entry new/old lists of
                        cube3.py:183-256, cross-checked against the literals in cpp/environments.h:75-105
  puzzle_tables.json    swap_zero_idxs for dim 4..7 (n_puzzle.py:174-214)
  cube3_cfg1.npz        BASELINE config 1: seeded generate_states(10000,(0,26)) parents, sha256 of the
                        12 children of every parent, is_solved flags, nnet input digest, and the full
                        children of the first 256 parents
  puzzle_cfg1.npz       same for puzzle15 / puzzle48 (2000 seeded states each)
  paths_<env>.npz       every (state, action, next_state) triple of results/<env>/results.pkl
  optimal_<env>.npz     data/<env>/test start states + optimal solutions (move indices)
  nnet_<env>.npz        reference ResnetModel (trained weights) cost-to-go for 256 states (fp32, CPU)
"""
import hashlib
import json
import os
import pickle
import random
import re
import sys

import numpy as np

np.float = float  # numpy>=1.24 shim for the reference
np.int = int
REF = os.environ.get("DCB_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
OUT = os.path.dirname(os.path.abspath(__file__))

import torch  # noqa: E402
from environments.cube3 import Cube3  # noqa: E402
from environments.n_puzzle import NPuzzle  # noqa: E402
from utils import nnet_utils  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cpp_literal_tables():
    src = open(os.path.join(REF, "cpp/environments.h")).read()
    out = {}
    for name in ("rotateIdxs_old", "rotateIdxs_new"):
        m = re.search(name + r"\[12\]\[24\]\s*=\s*\{(.*?)\};", src, re.S)
        rows = re.findall(r"\{([0-9,\s]+)\}", m.group(1))
        out[name] = [[int(x) for x in r.split(",")] for r in rows]
        assert len(out[name]) == 12 and all(len(r) == 24 for r in out[name])
    return out


def states_np(states, attr):
    return np.stack([np.asarray(getattr(s, attr)) for s in states]).astype(np.uint8)


def expand_ref(env, states, attr):
    exp, _ = env.expand(states)
    ch = np.stack([np.stack([np.asarray(getattr(c, attr)) for c in row]) for row in exp]).astype(np.uint8)
    flat = [c for row in exp for c in row]
    solved = env.is_solved(flat).reshape(len(states), -1)
    return ch, solved


def main():
    # ---- tables -------------------------------------------------------------------------------
    cube = Cube3()
    perm = np.tile(np.arange(54), (12, 1))
    new_l, old_l = [], []
    for a, m in enumerate(cube.moves):
        perm[a, cube.rotate_idxs_new[m]] = cube.rotate_idxs_old[m]
        new_l.append([int(x) for x in cube.rotate_idxs_new[m]])
        old_l.append([int(x) for x in cube.rotate_idxs_old[m]])
    lit = cpp_literal_tables()
    assert lit["rotateIdxs_new"] == new_l and lit["rotateIdxs_old"] == old_l, "python tables != C++ literals"
    json.dump({"moves": cube.moves, "moves_rev": cube.moves_rev, "perm": perm.tolist(),
               "idxs_new": new_l, "idxs_old": old_l}, open(f"{OUT}/cube3_tables.json", "w"))
    pz = {}
    for dim in (4, 5, 6, 7):
        e = NPuzzle(dim)
        pz[str(dim)] = {"swap_zero_idxs": e.swap_zero_idxs.astype(int).tolist(),
                        "goal": e.goal_tiles.astype(int).tolist(), "moves": e.moves, "moves_rev": e.moves_rev}
    json.dump(pz, open(f"{OUT}/puzzle_tables.json", "w"))

    # ---- config 1: seeded scrambles, all moves -------------------------------------------------
    np.random.seed(0); random.seed(0)
    states, depths = cube.generate_states(10000, (0, 26))
    par = states_np(states, "colors")
    ch, solved = expand_ref(cube, states, "colors")
    nn_in = cube.state_to_nnet_input(states)[0]
    np.savez_compressed(f"{OUT}/cube3_cfg1.npz", parents=par, depths=np.array(depths, np.int16),
                        children_sha256=np.array(sha(ch)), solved=np.packbits(solved),
                        n_solved=np.int64(solved.sum()), nnet_in_sha256=np.array(sha(nn_in)),
                        children_head=ch[:256])
    for name, dim in (("puzzle15", 4), ("puzzle48", 7)):
        e = NPuzzle(dim)
        np.random.seed(1); random.seed(1)
        st, dep = e.generate_states(2000, (0, 60))
        p = states_np(st, "tiles")
        c, sv = expand_ref(e, st, "tiles")
        np.savez_compressed(f"{OUT}/{name}_cfg1.npz", parents=p, depths=np.array(dep, np.int16),
                            children_sha256=np.array(sha(c)), solved=np.packbits(sv),
                            n_solved=np.int64(sv.sum()), children_head=c[:256])

    # ---- golden (s, a, s') triples from the shipped BWAS results --------------------------------
    for name, attr in (("cube3", "colors"), ("puzzle15", "tiles"), ("puzzle48", "tiles")):
        r = pickle.load(open(f"{REF}/results/{name}/results.pkl", "rb"))
        flat, moves, offs = [], [], [0]
        for path, soln in zip(r["paths"], r["solutions"]):
            assert len(path) == len(soln) + 1
            flat.append(states_np(path, attr)); moves.extend(soln); offs.append(offs[-1] + len(path))
        np.savez_compressed(f"{OUT}/paths_{name}.npz", states=np.concatenate(flat),
                            moves=np.array(moves, np.uint8), offsets=np.array(offs, np.int64),
                            times=np.array(r["times"], np.float64),
                            num_nodes_generated=np.array(r["num_nodes_generated"], np.int64))

    # ---- optimal solutions shipped with the test data ------------------------------------------
    d = pickle.load(open(f"{REF}/data/cube3/test/data_0.pkl", "rb"))
    mv = {m: i for i, m in enumerate(cube.moves)}
    sol = [[mv["%s%i" % (f, n)] for f, n in s] for s in d["solutions"]]
    np.savez_compressed(f"{OUT}/optimal_cube3.npz", states=states_np(d["states"], "colors"),
                        moves=np.array([m for s in sol for m in s], np.uint8),
                        offsets=np.cumsum([0] + [len(s) for s in sol]).astype(np.int64))
    d = pickle.load(open(f"{REF}/data/puzzle15/test/data_0.pkl", "rb"))
    mv = {m: i for i, m in enumerate(NPuzzle.moves)}
    sol = [[mv[c] for c in s] for s in d["solutions"]]
    np.savez_compressed(f"{OUT}/optimal_puzzle15.npz", states=states_np(d["states"], "tiles"),
                        moves=np.array([m for s in sol for m in s], np.uint8),
                        offsets=np.cumsum([0] + [len(s) for s in sol]).astype(np.int64))
    d = pickle.load(open(f"{REF}/data/puzzle48/test/data_0.pkl", "rb"))
    np.savez_compressed(f"{OUT}/start_puzzle48.npz", states=states_np(d["states"], "tiles"))

    # ---- reference cost-to-go with the trained weights (fp32, CPU) -----------------------------
    torch.set_num_threads(8)
    for name, env, attr in (("cube3", cube, "colors"), ("puzzle15", NPuzzle(4), "tiles"), ("puzzle48", NPuzzle(7), "tiles")):
        np.random.seed(2); random.seed(2)
        st, _ = env.generate_states(255, (0, 30))
        st = st + env.generate_goal_states(1)
        nnet = nnet_utils.load_nnet(f"{REF}/saved_models/{name}/current/model_state_dict.pt", env.get_nnet_model(),
                                    device=torch.device("cpu"))
        hf = nnet_utils.get_heuristic_fn(nnet, torch.device("cpu"), env, clip_zero=False)
        with torch.no_grad():
            ctg = hf(st)
        np.savez_compressed(f"{OUT}/nnet_{name}.npz", states=states_np(st, attr), ctg=ctg.astype(np.float32))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
