"""One-off derivation: find a 3-D sticker embedding (per face: outward normal, i-axis, j-axis)
under which plain 90-degree layer rotations reproduce the reference's 12 cube3 move permutations
(tests/golden/cube3_perm.json, produced from the reference by tests/golden/make_golden.py).
The embedding found here is hard-coded in deepcubea_b200/environments/cube3_geometry.py, which
builds the permutation tables geometrically instead of from index lists.
"""
import itertools, json, sys
import numpy as np

perm_ref = np.array(json.load(open(sys.argv[1]))["perm"])  # perm[a][new] = old

AX = [np.array(v) for v in ([1,0,0],[-1,0,0],[0,1,0],[0,-1,0],[0,0,1],[0,0,-1])]

def rot90(axis, sign):
    # Rodrigues for +-90 degrees about a unit axis (integer matrix)
    a = np.array(axis); K = np.array([[0,-a[2],a[1]],[a[2],0,-a[0]],[-a[1],a[0],0]])
    return np.outer(a,a) + sign*K

def build(normals, frames, sense):
    pos = {}; idx_of = {}
    for f in range(6):
        n = normals[f]; u, v = frames[f]
        for i in range(3):
            for j in range(3):
                p = tuple(3*n + 2*(i-1)*u + 2*(j-1)*v)  # doubled coords: face plane at 3, cubies at -2,0,2
                pos[(f,i,j)] = np.array(p); idx_of[p] = 9*f+3*i+j
    perms = []
    for f in range(6):
        for s in (-1, 1):
            R = rot90(normals[f], s*sense)
            pm = np.arange(54)
            for (ff,i,j), p in pos.items():
                if p @ normals[f] >= 2:   # on the turning layer
                    q = tuple(R @ p)
                    pm[idx_of[q]] = 9*ff+3*i+j
            perms.append(pm)
    return np.array(perms)

found = []
for chir in (1, -1):
    normals = [np.array([0,0,1]), np.array([0,0,-1]), np.array([-1,0,0]), np.array([1,0,0]),
               np.array([0,chir,0]), np.array([0,-chir,0])]
    per_face = []
    for f in range(6):
        opts = []
        for u in AX:
            for v in AX:
                if u @ normals[f] == 0 and v @ normals[f] == 0 and u @ v == 0:
                    opts.append((u, v))
        per_face.append(opts)
    for sense in (1, -1):
        # prune: own-face 9 stickers of each move depend only on that face's frame
        ok_face = []
        for f in range(6):
            good = []
            for fr in per_face[f]:
                frames = [per_face[g][0] for g in range(6)]; frames[f] = fr
                pm = build(normals, frames, sense)
                if all((pm[2*f+k, 9*f:9*f+9] == perm_ref[2*f+k, 9*f:9*f+9]).all() for k in (0,1)):
                    good.append(fr)
            ok_face.append(good)
        for combo in itertools.product(*ok_face):
            pm = build(normals, list(combo), sense)
            if (pm == perm_ref).all():
                found.append((chir, sense, combo))
print(len(found), "embeddings reproduce the reference tables")
chir, sense, combo = found[0]
print("chirality", chir, "sense", sense)
for f, (u, v) in enumerate(combo):
    print(f, "UDLRBF"[f], "u=", u.tolist(), "v=", v.tolist())
