"""CPU oracle for the DeepCubeA hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (`deepcubea_b200/`) may import, link or execute this directory; only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` do.
Every function cites the reference file:line it restates.  The oracle is pinned against the fixtures in
`tests/golden/` that `tests/golden/make_golden.py` produced by running the unmodified reference.
"""
