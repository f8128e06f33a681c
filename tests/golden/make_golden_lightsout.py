#!/usr/bin/env python
"""Golden fixtures for Lights Out 7x7 from the UNMODIFIED reference (build container only):
    python tests/golden/make_golden_lightsout.py
  lightsout_tables.json   move_matrix[49][5] (environments/lights_out.py:31-42 == cpp/environments.cpp:133-155)
  lightsout7_cfg1.npz     seeded generate_states(2000,(0,50)) parents, sha256 of all 49 children, solved flags, heads
  paths_lightsout7.npz    every (s, a, s') triple of results/lightsout7/results.pkl (+ times, nodes)
  nnet_lightsout7.npz     reference ResnetModel cost-to-go (trained weights, CPU fp32) for 256 states
"""
import hashlib
import json
import os
import pickle
import random
import sys

import numpy as np

np.float = float
np.int = int
REF = os.environ.get("DCB_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
OUT = os.path.dirname(os.path.abspath(__file__))
import torch  # noqa: E402
from environments.lights_out import LightsOut  # noqa: E402
from utils import nnet_utils  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    env = LightsOut(7)
    json.dump({"move_matrix": env.move_matrix.astype(int).tolist()}, open(f"{OUT}/lightsout_tables.json", "w"))
    np.random.seed(4); random.seed(4)
    st, dep = env.generate_states(2000, (0, 50))
    par = np.stack([s.tiles for s in st]).astype(np.uint8)
    exp, _ = env.expand(st)
    ch = np.stack([np.stack([c.tiles for c in row]) for row in exp]).astype(np.uint8)
    solved = env.is_solved([c for row in exp for c in row]).reshape(2000, 49)
    np.savez_compressed(f"{OUT}/lightsout7_cfg1.npz", parents=par, depths=np.array(dep, np.int16), children_sha256=np.array(sha(ch)),
                        solved=np.packbits(solved), n_solved=np.int64(solved.sum()), children_head=ch[:64])
    r = pickle.load(open(f"{REF}/results/lightsout7/results.pkl", "rb"))
    flat, moves, offs = [], [], [0]
    for path, soln in zip(r["paths"], r["solutions"]):
        flat.append(np.stack([s.tiles for s in path]).astype(np.uint8)); moves.extend(soln); offs.append(offs[-1] + len(path))
    np.savez_compressed(f"{OUT}/paths_lightsout7.npz", states=np.concatenate(flat), moves=np.array(moves, np.uint8),
                        offsets=np.array(offs, np.int64), times=np.array(r["times"], np.float64),
                        num_nodes_generated=np.array(r["num_nodes_generated"], np.int64))
    torch.set_num_threads(8)
    np.random.seed(2); random.seed(2)
    s2, _ = env.generate_states(255, (0, 30))
    s2 = s2 + env.generate_goal_states(1)
    nnet = nnet_utils.load_nnet(f"{REF}/saved_models/lightsout7/current/model_state_dict.pt", env.get_nnet_model(), device=torch.device("cpu"))
    hf = nnet_utils.get_heuristic_fn(nnet, torch.device("cpu"), env, clip_zero=False)
    with torch.no_grad():
        ctg = hf(s2)
    np.savez_compressed(f"{OUT}/nnet_lightsout7.npz", states=np.stack([s.tiles for s in s2]).astype(np.uint8), ctg=ctg.astype(np.float32))
    print("lightsout fixtures written")


if __name__ == "__main__":
    main()
