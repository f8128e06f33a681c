"""Alias of deepcubea_b200.environments.cube3 (reference import path, used by pickles)."""
from deepcubea_b200.environments.cube3 import *  # noqa: F401,F403
from deepcubea_b200.environments.cube3 import Cube3, Cube3State  # noqa: F401
