#!/usr/bin/env python
"""Golden fixtures for the 4x4x4 cube from the UNMODIFIED reference (build container only).  The reference
has Cube4 only as a C++ class (cpp/environments.cpp:262-370); `make -C oracle ref` compiles it where it lies together with
oracle/ref_cube4_dump.cpp (ours), and this script runs that binary:
    python tests/golden/make_golden_cube4.py
  cube4_tables.json   perm[24][96]: the reference's children of the identity state (child[j] = parent[perm[a][j]])
  cube4_cfg1.npz      1500 seeded scrambles (depth 0..14) + 64 states around non-identity solved states: sha256 of all 24
                      children, Cube4::isSolved of every parent and child, first 64 parents' children verbatim
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(ROOT, "oracle", "_ref", "cube4_dump")


def reference_expand(states: np.ndarray):
    inp = "%d\n" % len(states) + "\n".join(" ".join(str(int(v)) for v in s) for s in states) + "\n"
    out = subprocess.run([BIN], input=inp, capture_output=True, text=True, check=True).stdout.split("\n")
    ch = np.zeros((len(states), 24, 96), np.uint8)
    solved = np.zeros((len(states), 25), np.uint8)
    for k in range(len(states)):
        blk = out[25 * k:25 * k + 25]
        ch[k] = np.array([[int(x) for x in ln.split()] for ln in blk[:24]], np.uint8)
        solved[k] = np.array([int(x) for x in blk[24].split()], np.uint8)
    return ch, solved


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    ident = np.arange(96, dtype=np.uint8)[None, :]
    ch0, s0 = reference_expand(ident)
    perm = ch0[0].astype(int)
    assert s0[0, 0] == 1
    json.dump({"perm": perm.tolist()}, open(f"{OUT}/cube4_tables.json", "w"))
    rng = np.random.RandomState(11)
    n = 1500
    depth = rng.randint(0, 15, size=n)
    st = np.repeat(ident, n, axis=0)
    for k in range(n):
        for a in rng.randint(0, 24, size=depth[k]):
            st[k] = st[k][perm[a]]
    # solved states other than the identity: whole-cube rotations (all four layers of an axis turned the same way) and
    # stickers exchanged inside a face (isSolved only looks at state[i] / 16, environments.cpp:356-366) -- and their neighbours
    extra = []
    for (a, b, c, d) in ((1, 13, 14, 2), (5, 17, 18, 6), (9, 21, 22, 10)):      # e.g. U0+1 U1+1 D1-1 D0-1 = a rotation about z
        s = ident[0].copy()
        for _ in range(rng.randint(1, 4)):
            for m in (a, b, c, d):
                s = s[perm[m]]
        extra.append(s)
    for _ in range(13):
        s = ident[0].copy()
        f = rng.randint(0, 6)
        i, j = rng.choice(16, 2, replace=False)
        s[16 * f + i], s[16 * f + j] = s[16 * f + j], s[16 * f + i]
        extra.append(s)
    near = []
    for s in extra:
        for a in rng.randint(0, 24, size=3):
            near.append(s[perm[a]])
    parents = np.concatenate([st, np.array(extra, np.uint8), np.array(near, np.uint8)])
    ch, solved = reference_expand(parents)
    np.savez_compressed(f"{OUT}/cube4_cfg1.npz", parents=parents, depths=depth.astype(np.int16),
                        children_sha256=np.array(hashlib.sha256(ch.tobytes()).hexdigest()), solved=solved,
                        children_head=ch[:64], n_solved=np.int64(solved.sum()))
    print("cube4 fixtures written:", parents.shape, "solved flags set:", int(solved.sum()))


if __name__ == "__main__":
    sys.exit(main())
