"""Alias of deepcubea_b200.environments.n_puzzle (reference import path, used by pickles)."""
from deepcubea_b200.environments.n_puzzle import *  # noqa: F401,F403
from deepcubea_b200.environments.n_puzzle import NPuzzle, NPuzzleState  # noqa: F401
