#!/usr/bin/env python
"""Copy the reference's trained cost-to-go weights and test start states into assets/ (git-ignored; the
directory travels to the GPU box with the working tree).  Data only -- no reference source is copied.
bench.py and the GPU tests use them when present and fall back to seeded random-init weights otherwise.

    python tools/fetch_assets.py [env ...]        (default: the BASELINE.json environments cube3 puzzle15 puzzle48, and lightsout7)
"""
import os
import shutil
import sys

REF = os.environ.get("DCB_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fetch(envs):
    if not os.path.isdir(REF):
        print("reference not present; assets unchanged")
        return
    for env in envs:
        for rel in ("saved_models/%s/current/model_state_dict.pt" % env, "data/%s/test/data_0.pkl" % env):
            src, dst = os.path.join(REF, rel), os.path.join(ROOT, "assets", rel)
            if os.path.exists(src) and not (os.path.exists(dst) and os.path.getsize(dst) == os.path.getsize(src)):
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                print("copied", rel)


if __name__ == "__main__":
    fetch(sys.argv[1:] or ["cube3", "puzzle15", "puzzle48", "lightsout7"])
